#!/bin/bash
# final profiles of round 2: launch lists (bench command, cluster call) + ncu --set full of the mel kernel (both value types)
# and of the filter GEMM
mkdir -p gpurun_out
bash scripts/gpu_mel_exp.sh > /dev/null 2>&1; cat gpurun_out/mel_exp.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --only-main > gpurun_out/ncu_bench_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mel512 -s 2 -c 1 -f -o gpurun_out/prof_mel_f32 python scripts/profile_target.py mel32 3 > gpurun_out/ncu_mel32_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mel512 -s 2 -c 1 -f -o gpurun_out/prof_mel python scripts/profile_target.py mel 3 > gpurun_out/ncu_mel_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_cluster.csv python scripts/profile_target.py cluster 2 > gpurun_out/ncu_cluster_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile128 -s 1 -c 1 -f -o gpurun_out/prof_ahc_filter python scripts/profile_target.py cluster 2 > gpurun_out/ncu_filter_full.log 2>&1
ls -la gpurun_out | head -40
