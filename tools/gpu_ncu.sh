#!/bin/bash
# ncu: launch list (per-kernel durations) + full capture of the blend kernels. 1 GPU only.
mkdir -p gpurun_out
WL=${1:-cfg3_1M_sh3_1080p}
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$WL.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --workload $WL > gpurun_out/ncu_bench_$WL.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:blend_ -s 6 -c 2 -f -o gpurun_out/prof_blend_$WL \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --workload $WL > gpurun_out/ncu_full_$WL.log 2>&1
ls -la gpurun_out/
