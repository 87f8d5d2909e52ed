#!/bin/bash
bash scripts/gpu_tests_only.sh -k "mel or audio or swift"
bash scripts/gpu_mel_exp.sh ncu
python scripts/ncu_lines.py gpurun_out/prof_mel_f32.ncu-rep 360001 12 | head -16
