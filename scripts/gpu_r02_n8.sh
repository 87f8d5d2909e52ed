#!/bin/bash
# bench at N = 8 (torchrun, NCCL): the multi-GPU configurations of BASELINE (configs[3], configs[4]) + weak-scaling main line
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -14
N=${1:-8}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 20 --warmup 5 2>gpurun_out/bench$N.err > gpurun_out/bench$N.json; tail -c 800 gpurun_out/bench$N.err
python - <<PY
import json
n = $N
try:
    d = json.loads([l for l in open(f'gpurun_out/bench{n}.json') if l.startswith('{')][-1]); c = d['cluster']
    print(f"N={n} mel {d['value']:.0f} ({d['ms_per_step']:.4f} ms) f64 {d['f64_transform']['ms_per_step']:.4f} e2e {d['e2e']['value']:.1f} ({d['e2e']['ms_per_step']:.3f} ms, floor {d['e2e']['copy_floor_ms']:.3f}) i16 {d['e2e_i16']['value']:.1f} ({d['e2e_i16']['ms_per_step']:.3f}, floor {d['e2e_i16']['copy_floor_ms']:.3f})")
    print('   parity', d['parity'])
    print('   cluster', round(c['value']), c['ms_per_step'], c['stages_ms'], c.get('labels_equal_ref'), c.get('labels_deterministic_all_ranks'))
    print('   c4', d['c4']['e2e'], d['c4']['clips_equal'], d['c4']['clips_recomputed_on_rank0'])
    print('   c5', d['c5']['e2e'], d['c5']['labels_equal_ref'], d['c5']['ahc_ms_per_meeting'])
    print('   clocks', d['clocks'], d['host_binding'])
except Exception as e:
    print('parse failed', e)
PY
