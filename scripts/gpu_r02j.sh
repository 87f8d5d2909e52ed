#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_tests_only.sh -k "mel or audio"
bash scripts/gpu_mel_exp.sh 2>&1 | head -4
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mel512 -s 2 -c 1 -f -o gpurun_out/prof_mel_f32 python scripts/profile_target.py mel32 3 > gpurun_out/ncu_mel32_full.log 2>&1
python scripts/ncu_lines.py gpurun_out/prof_mel_f32.ncu-rep 360001 6 | head -3
