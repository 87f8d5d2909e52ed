#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_tests_only.sh
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench1.err > gpurun_out/bench1.json; tail -c 300 gpurun_out/bench1.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/bench1.json') if l.startswith('{')][-1]); c = d['cluster']
print(f"mel {d['value']:.0f} ({d['ms_per_step']:.4f} ms) sustained {d['sustained']['ms_per_step']:.4f} f64 {d['f64_transform']['ms_per_step']:.4f} e2e {d['e2e']['ms_per_step']:.3f} i16 {d['e2e_i16']['ms_per_step']:.3f}")
print('cluster', round(c['value']), c['ms_per_step'], c['stages_ms'], c.get('labels_equal_ref'), c.get('labels_equal_cpu'))
print('c4', d['c4']['e2e']['value'], 'c5', d['c5']['e2e'], d['c5']['labels_equal_ref'])
print('streaming', {k: v['p50_us'] for k, v in d['streaming'].items() if k != 'api'})
PY
