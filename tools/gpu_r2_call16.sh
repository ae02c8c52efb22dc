#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -q -m gpu 2>&1 | tail -8 > gpurun_out/c16_pytest.log
for wl in cfg4_5M_sh3_1080p cfg5_city_16k_540p; do
    timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --workload $wl 2>gpurun_out/c16_bench_$wl.err | tail -1 > gpurun_out/c16_bench_$wl.json
done
timeout 200 python tools/gpu_ncu_striped.py --world 8 --rank 3 --steps 5 > gpurun_out/c16_striped_rank3of8.log 2>&1
