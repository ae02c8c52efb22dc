#!/bin/bash
mkdir -p gpurun_out
timeout 250 compute-sanitizer --tool memcheck --print-limit 8 python tools/gpu_sanitize.py > gpurun_out/c9_memcheck.log 2>&1
timeout 250 compute-sanitizer --tool racecheck --print-limit 8 python tools/gpu_sanitize.py > gpurun_out/c9_racecheck.log 2>&1
timeout 900 python -m pytest tests/ -q -m gpu 2>&1 | tail -25 > gpurun_out/c9_pytest.log
for wl in cfg4_5M_sh3_1080p cfg5_city_16k_540p cfg3_1M_sh3_1080p; do
    timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --workload $wl 2>gpurun_out/c9_bench_$wl.err | tail -1 > gpurun_out/c9_bench_$wl.json
done
