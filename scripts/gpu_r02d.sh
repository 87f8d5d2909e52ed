#!/bin/bash
# full GPU test-suite + bench at N = 1 and N = 2 (torchrun, NCCL)
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -12
bash scripts/gpu_tests_only.sh
timeout 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench1.err > gpurun_out/bench1.json; tail -c 300 gpurun_out/bench1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 2>gpurun_out/bench2.err > gpurun_out/bench2.json; tail -c 600 gpurun_out/bench2.err
python - <<'PY'
import json
for n in (1, 2):
    try:
        txt = [l for l in open(f'gpurun_out/bench{n}.json') if l.startswith('{')][-1]
        d = json.loads(txt); c = d['cluster']
        print(f"N={n} mel {d['value']:.0f} ({d['ms_per_step']:.4f} ms) f64 {d['f64_transform']['ms_per_step']:.4f} e2e {d['e2e']['value']:.1f} ({d['e2e']['ms_per_step']:.3f} ms, floor {d['e2e']['copy_floor_ms']:.3f}) i16 {d['e2e_i16']['value']:.1f} ({d['e2e_i16']['ms_per_step']:.3f}, floor {d['e2e_i16']['copy_floor_ms']:.3f})")
        print('   parity', d['parity'])
        print('   cluster', round(c['value']), c['ms_per_step'], c['stages_ms'], c.get('labels_equal_ref'), c.get('labels_deterministic_all_ranks'))
        print('   c4', d['c4']['e2e'], d['c4']['clips_equal'], d['c4']['clips_recomputed_on_rank0'])
        print('   c5', d['c5']['e2e'], d['c5']['labels_equal_ref'], d['c5']['ahc_ms_per_meeting'])
        print('   streaming', d.get('streaming')); print('   clocks', d['clocks']); print('   cpu', d.get('cpu_baseline'))
    except Exception as e:
        print('N', n, 'parse failed', e)
PY
