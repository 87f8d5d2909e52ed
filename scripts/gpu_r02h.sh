#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_tests_only.sh -k "vbx or cluster or pipeline or batch or constrained or export or kmeans or speaker or next_rows"
timeout 600 python bench.py --workload cluster --steps 5 --warmup 3 2>gpurun_out/benchc.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('cluster', round(d['value']), d['ms_per_step'], d['stages_ms'], d.get('labels_equal_ref'), d.get('labels_equal_cpu'))"
tail -c 300 gpurun_out/benchc.err
