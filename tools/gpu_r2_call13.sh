#!/bin/bash
mkdir -p gpurun_out
timeout 250 compute-sanitizer --tool memcheck --print-limit 8 python tools/gpu_sanitize.py 2>&1 | tail -6 > gpurun_out/c13_memcheck.log
timeout 250 compute-sanitizer --tool racecheck --print-limit 4 python tools/gpu_sanitize.py 2>&1 | grep -E "RACECHECK SUMMARY|Race reported|and (Read|Write)|^ok" | cut -c1-220 | sort | uniq -c | sort -rn | head -12 > gpurun_out/c13_racecheck.log
timeout 900 python -m pytest tests/ -q -m gpu 2>&1 | tail -25 > gpurun_out/c13_pytest.log
for wl in cfg5_city_16k_540p cfg2_100k_sh0_512 cfg4_5M_sh3_1080p; do
  for impl in ours reference; do
    timeout 300 python bench.py --impl $impl --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --workload $wl 2>gpurun_out/c13_bench_${impl}_$wl.err | tail -1 > gpurun_out/c13_bench_${impl}_$wl.json
  done
done
for arm in reference ours ours_wrapper ours_fused; do
  timeout 200 python tools/config5_gstep.py --arm $arm --steps 100 --warmup 10 --profile > gpurun_out/c13_cfg5_$arm.json 2>gpurun_out/c13_cfg5_$arm.err
done
