#!/bin/bash
mkdir -p gpurun_out
timeout 250 compute-sanitizer --tool memcheck --print-limit 8 python tools/gpu_sanitize.py 2>&1 | tail -5 > gpurun_out/c15_memcheck.log
timeout 900 python -m pytest tests/ -q -m gpu 2>&1 | tail -25 > gpurun_out/c15_pytest.log
for wl in cfg4_5M_sh3_1080p cfg5_city_16k_540p; do
    timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --workload $wl 2>gpurun_out/c15_bench_$wl.err | tail -1 > gpurun_out/c15_bench_$wl.json
done
timeout 200 python tools/gpu_ncu_striped.py --world 8 --rank 3 --steps 5 > gpurun_out/c15_striped_rank3of8.log 2>&1
timeout 200 python tools/gpu_ncu_striped.py --world 2 --rank 1 --steps 5 > gpurun_out/c15_striped_rank1of2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on \
  -k regex:'project_kernel|stripe_select_kernel|preprocess_bwd_kernel|onesweep_pass_kernel|emit_scan' \
  -s 20 -c 10 -o gpurun_out/r02_full_striped_rank3of8 python tools/gpu_ncu_striped.py --world 8 --rank 3 --steps 1 > gpurun_out/c15_ncu_striped.log 2>&1
