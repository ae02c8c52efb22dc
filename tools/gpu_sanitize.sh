#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 280 compute-sanitizer --tool $tool --print-limit 5 python tools/gpu_sanitize.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|errors' gpurun_out/sanitize_$tool.log | tail -2 | tr '\n' ' ')  $(grep -c '^ok' gpurun_out/sanitize_$tool.log) workloads ok"
done
