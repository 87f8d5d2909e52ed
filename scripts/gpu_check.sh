#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 300 python scripts/gpu_first_light.py ahc 2>&1 | grep -E "N=10000|N=1000 |N=3000|dups|lattice" 
timeout 600 python bench.py --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('mel value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'])
c=d['cluster']; print('cluster value',c['value'],'ms',c['ms_per_step'],'e2e',c['e2e']['value'],c['stages_ms'])"
