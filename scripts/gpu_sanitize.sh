#!/bin/bash
# compute-sanitizer passes over a small end-to-end run; summaries under gpurun_out/
mkdir -p gpurun_out
for tool in ${TOOLS:-memcheck racecheck synccheck}; do
  FA_AHC_FILTER_MIN_N=2 timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_target.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool: $(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/sanitize_$tool.log | tail -1) | $(grep -c "sanitize target done" gpurun_out/sanitize_$tool.log) run(s) completed"
done
