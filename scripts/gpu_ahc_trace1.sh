#!/bin/bash
# one-size trace run: FLAGS="4 12" N=10000 bash scripts/gpu_ahc_trace1.sh
for f in ${FLAGS:-4}; do
  echo "FA_AHC_FLAGS=$f"
  FA_AHC_FLAGS=$f NPTS=${N:-10000} timeout 120 python - <<'PY' 2>&1 | grep -E "trace|N=|Error|error"
import sys, os; sys.path.insert(0,'.')
import numpy as np
from fluidaudio_b200 import synth, _lib, clustering as cl
N = int(os.environ["NPTS"])
emb,_ = synth.speaker_embeddings(N,256,8,seed=42)
x = emb.astype(np.float64); x /= np.linalg.norm(x,axis=1,keepdims=True)
for rep in range(2):
    st,z = cl.centroid_linkage(x)
ms=np.zeros(4,np.float32); _lib.load().fa_ahc_last_stage_ms(ms.ctypes.data); print("N=",N,st,ms, flush=True)
PY
done
