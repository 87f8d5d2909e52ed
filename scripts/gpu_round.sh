#!/bin/bash
# One GPU session: tests, smoke, bench, ncu launch lists and full captures.  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 600 python bench.py 2>gpurun_out/bench.err > gpurun_out/bench.json; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); c=d['cluster']
print('mel', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'cpu', d.get('cpu_baseline',{}).get('value'))
print('cluster', c['value'], c['ms_per_step'], 'e2e', c['e2e']['value'], 'cpu', c.get('cpu_baseline',{}).get('value'), c['stages_ms'])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>>gpurun_out/bench.err > gpurun_out/bench_reference.json
if [ "$1" == "ncu" ]; then
# launch list of the bench command itself (numbers printed under ncu are not bench values)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_list.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_mel.csv python scripts/profile_target.py mel 4 > gpurun_out/ncu_mel_list.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cluster.csv python scripts/profile_target.py cluster 2 > gpurun_out/ncu_cluster_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mel512 -s 2 -c 1 -f -o gpurun_out/prof_mel python scripts/profile_target.py mel 3 > gpurun_out/ncu_mel_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ahc_merge|ahc_init_nn" -c 2 -f -o gpurun_out/prof_ahc python scripts/profile_target.py cluster 1 > gpurun_out/ncu_ahc_full.log 2>&1
fi
ls gpurun_out
