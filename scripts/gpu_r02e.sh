#!/bin/bash
# profiles for round 2: mel experiments, tensor-core filterbank probe, launch list of the bench command, full ncu captures
mkdir -p gpurun_out
bash scripts/gpu_tests_only.sh -k "mel or audio"
bash scripts/gpu_mel_exp.sh
timeout 600 python scripts/gpu_tc_filterbank_probe.py 2>&1 | tail -12 | tee gpurun_out/tc_filterbank_probe.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --only-main > gpurun_out/ncu_bench_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mel512 -s 2 -c 1 -f -o gpurun_out/prof_mel_f32 python scripts/profile_target.py mel32 3 > gpurun_out/ncu_mel32_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mel512 -s 2 -c 1 -f -o gpurun_out/prof_mel python scripts/profile_target.py mel 3 > gpurun_out/ncu_mel_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_cluster.csv python scripts/profile_target.py cluster 2 > gpurun_out/ncu_cluster_list.log 2>&1
ls gpurun_out | head -40
