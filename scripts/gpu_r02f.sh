#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_tests_only.sh -k "linkage or ahc or cluster or baseline or pipeline or batch"
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/ahc_filter.log
import os, sys, time, subprocess, numpy as np
sys.path.insert(0, '.')
code = r"""
import sys, numpy as np, ctypes as C
sys.path.insert(0, '.')
from fluidaudio_b200 import _lib, synth
from fluidaudio_b200.clustering import OfflineClusterer, centroid_linkage
from oracle import oracle as O
L = _lib.load()
for n, k, w, seed in ((10000, 8, None, 42), (5000, 4, (0.4, 0.3, 0.2, 0.1), 0)):
    emb, _ = synth.speaker_embeddings(n, 256, k, weights=w, seed=seed)
    x = O.l2_normalize_rows(emb.astype(np.float64))
    for _ in range(3):
        st, z = centroid_linkage(x)
    ms = (C.c_float * 4)(); L.fa_ahc_last_stage_ms(ms)
    print(f"N={n}: init {ms[0]:.3f} ms  host {ms[1]:.3f}  merge {ms[2]:.3f}  total {ms[3]:.3f}", flush=True)
"""
for env in ({"FA_AHC_FILTER_MIN_N": "0"}, {}):
    print("env", env, flush=True)
    print(subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True).stdout)
PY
timeout 600 python scripts/gpu_tc_filterbank_probe.py 2>&1 | tail -9 | tee gpurun_out/tc_filterbank_probe.log
