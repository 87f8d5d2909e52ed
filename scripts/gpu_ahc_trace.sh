#!/bin/bash
for f in ${FLAGS:-4 8}; do
  echo "FA_AHC_FLAGS=$f"
  FA_AHC_FLAGS=$f timeout 120 python - <<'PY' 2>&1 | grep -E "trace|N=|Error|error" 
import sys; sys.path.insert(0,'.')
import numpy as np, time
from fluidaudio_b200 import synth, _lib, clustering as cl
from oracle import oracle as O
for N in (2, 3, 4, 100, 1000, 5000, 10000):
    emb,_ = synth.speaker_embeddings(N,256,8,seed=42)
    x = O.l2_normalize_rows(emb.astype(np.float64))
    for rep in range(2):
        st,z = cl.centroid_linkage(x)
    st2,z2 = O.centroid_linkage(x) if N<=5000 else (0,z)
    ms=np.zeros(4,np.float32); _lib.load().fa_ahc_last_stage_ms(ms.ctypes.data); print("N=",N,st,ms,"bit-exact",np.array_equal(z,z2), flush=True)
PY
done
