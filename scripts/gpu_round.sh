#!/bin/bash
# One GPU session: tests, smoke, bench, ncu launch lists and full captures.  Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 600 python bench.py 2>gpurun_out/bench.err | tee gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>>gpurun_out/bench.err | tee gpurun_out/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_mel.csv python scripts/profile_target.py mel 4 > gpurun_out/ncu_mel_list.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cluster.csv python scripts/profile_target.py cluster 2 > gpurun_out/ncu_cluster_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mel512 -s 2 -c 1 -f -o gpurun_out/prof_mel python scripts/profile_target.py mel 3 > gpurun_out/ncu_mel_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ahc_merge|ahc_init_nn" -c 2 -f -o gpurun_out/prof_ahc python scripts/profile_target.py cluster 1 > gpurun_out/ncu_ahc_full.log 2>&1
ls -la gpurun_out
