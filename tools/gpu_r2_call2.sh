#!/bin/bash
# round-2 call 2: restructured pipeline -- sanitizer on tiny cases first, then the gpu suite, then bench
mkdir -p gpurun_out
timeout 200 python tools/gpu_sanitize.py > gpurun_out/c2_plain.log 2>&1; echo "plain rc=$?" >> gpurun_out/c2_plain.log
timeout 280 compute-sanitizer --tool memcheck --print-limit 8 python tools/gpu_sanitize.py > gpurun_out/c2_memcheck.log 2>&1
timeout 900 python -m pytest tests/ -q -m gpu 2>&1 | tail -60 > gpurun_out/c2_pytest.log
for wl in cfg4_5M_sh3_1080p cfg3_1M_sh3_1080p cfg5_city_16k_540p; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --workload $wl 2>gpurun_out/c2_bench_$wl.err | tail -1 > gpurun_out/c2_bench_$wl.json
done
