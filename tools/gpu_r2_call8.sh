#!/bin/bash
# 1 GPU: full gpu suite, benches (headline, 1M, 16k with native binding), config-5 step profile, ncu launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -q -m gpu 2>&1 | tail -15 > gpurun_out/c8_pytest.log
for wl in cfg4_5M_sh3_1080p cfg5_city_16k_540p; do
  for impl in ours reference; do
    timeout 300 python bench.py --impl $impl --steps 30 --warmup 5 --no-cpu-baseline --workload $wl 2>gpurun_out/c8_bench_${impl}_$wl.err | tail -1 > gpurun_out/c8_bench_${impl}_$wl.json
  done
done
for arm in reference ours ours_wrapper; do
  timeout 200 python tools/config5_gstep.py --arm $arm --steps 50 --warmup 10 --profile > gpurun_out/c8_cfg5_$arm.json 2>gpurun_out/c8_cfg5_$arm.err
done
GCR_HOST_TIMING=1 timeout 100 python tools/gpu_overhead.py cfg5_city_16k_540p 2>&1 | tail -12 > gpurun_out/c8_overhead.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/r02_launches_cfg4_5M.csv \
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/c8_ncu_launch.log 2>&1
