#!/bin/bash
# Round-2 final single-GPU record: full gpu suite, sanitizer, every BASELINE config on both arms,
# config-5 harness (4 arms), ncu launch list + one --set full pass, striped-rank stage profile.
mkdir -p gpurun_out
O=gpurun_out/final
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > ${O}_gpu.txt; nproc >> ${O}_gpu.txt
timeout 900 python -m pytest tests/ -q -m gpu 2>&1 | tail -8 > ${O}_pytest.log
timeout 250 compute-sanitizer --tool memcheck --print-limit 8 python tools/gpu_sanitize.py 2>&1 | grep -E "^ok|ERROR SUMMARY" > ${O}_memcheck.log
timeout 250 compute-sanitizer --tool racecheck --print-limit 8 python tools/gpu_sanitize.py 2>&1 | grep -E "^ok|RACECHECK SUMMARY|Race reported" | cut -c1-200 > ${O}_racecheck.log
timeout 250 compute-sanitizer --tool synccheck --print-limit 8 python tools/gpu_sanitize.py 2>&1 | grep -E "^ok|ERROR SUMMARY" > ${O}_synccheck.log
for impl in reference ours; do
  timeout 400 python bench.py --impl $impl --steps 30 --warmup 5 2>${O}_bench_${impl}_cfg4_5M_sh3_1080p.err | tail -1 > ${O}_bench_${impl}_cfg4_5M_sh3_1080p.json
  for wl in cfg3_1M_sh3_1080p cfg2_100k_sh0_512 city_5M_precomp_1080p cfg5_city_16k_540p city_500k_540p; do
    timeout 300 python bench.py --impl $impl --steps 50 --warmup 5 --no-cpu-baseline --workload $wl 2>${O}_bench_${impl}_$wl.err | tail -1 > ${O}_bench_${impl}_$wl.json
  done
done
for arm in reference ours ours_wrapper ours_fused; do
  timeout 200 python tools/config5_gstep.py --arm $arm --steps 200 --warmup 20 > ${O}_cfg5_${arm}_n1.json 2>${O}_cfg5_${arm}_n1.err
done
timeout 200 python tools/gpu_ncu_striped.py --world 8 --rank 3 --steps 5 > ${O}_striped_rank3of8.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/r02_launches_cfg4_5M.csv \
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > ${O}_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:'blend_bwd_kernel|blend_fwd_kernel|project_kernel|preprocess_bwd_kernel|emit_scan_kernel|onesweep_pass_kernel|radix_hist|tile_ranges' \
  -s 45 -c 15 -o gpurun_out/r02_full_cfg4_5M python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > ${O}_ncu_full.log 2>&1
