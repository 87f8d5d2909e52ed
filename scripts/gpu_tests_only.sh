#!/bin/bash
# GPU test-suite only (all tests, no early exit); full log in gpurun_out/pytest_gpu_full.log
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q "$@" > gpurun_out/pytest_gpu_full.log 2>&1
tail -15 gpurun_out/pytest_gpu_full.log | tee gpurun_out/pytest_gpu.log
