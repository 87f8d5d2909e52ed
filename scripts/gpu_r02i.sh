#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_tests_only.sh
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
FA_AHC_FILTER_IMPL=0 timeout 600 python scripts/gpu_ahc_filter_ab.py child 2>&1 | tail -1 | tee gpurun_out/ahc_filter_ab2.txt
