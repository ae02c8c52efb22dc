#!/bin/bash
# 1 GPU: new tests, launch list, one ncu --set full pass over every kernel of one fwd+bwd step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -q -m gpu 2>&1 | tail -25 > gpurun_out/c10_pytest.log
timeout 250 compute-sanitizer --tool racecheck --print-limit 4 python tools/gpu_sanitize.py 2>&1 | grep -E "RACECHECK SUMMARY|Race reported|and (Read|Write)" | cut -c1-220 | sort | uniq -c | sort -rn | head -20 > gpurun_out/c10_racecheck.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/r02_launches_cfg4_5M.csv \
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/c10_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:'blend_bwd_kernel|blend_fwd_kernel|project_kernel|preprocess_bwd_kernel|emit_scan_kernel|onesweep_pass_kernel|radix_hist|tile_ranges' \
  -s 45 -c 15 -o gpurun_out/r02_full_cfg4_5M python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/c10_ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep >> gpurun_out/c10_ncu_full.log
for wl in cfg4_5M_sh3_1080p cfg5_city_16k_540p; do
    timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --workload $wl 2>gpurun_out/c10_bench_$wl.err | tail -1 > gpurun_out/c10_bench_$wl.json
done
