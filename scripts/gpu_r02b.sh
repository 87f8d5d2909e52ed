#!/bin/bash
bash scripts/gpu_tests_only.sh
bash scripts/gpu_mel_exp.sh ncu
timeout 600 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench.err > gpurun_out/bench.json; tail -c 400 gpurun_out/bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench.json')); c = d['cluster']
print('mel', d['ms_per_step'], 'f64', d['f64_transform']['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'i16', d['e2e_i16']['ms_per_step'])
print('cluster', c['ms_per_step'], c['stages_ms'], c.get('labels_equal_ref'), c.get('labels_equal_cpu'))
print('streaming', d.get('streaming'))
PY
