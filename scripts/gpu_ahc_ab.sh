#!/bin/bash
for f in 0 1 2 3; do
  echo "FA_AHC_FLAGS=$f"
  FA_AHC_FLAGS=$f timeout 300 python scripts/gpu_first_light.py ahc 2>&1 | grep -E "N=10000|N=1000 " 
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "batch_of_sets or reentrant or fallback" 2>&1 | tail -5
