#!/bin/bash
# full ncu capture of the three heaviest kernels at the headline workload (1 GPU only)
mkdir -p gpurun_out
WL=${1:-cfg4_5M_sh3_1080p}
ncu --set full --clock-control none --import-source on -k regex:"blend_|preprocess_bwd" -s 9 -c 3 -f -o gpurun_out/prof_r01_v2_$WL \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --workload $WL > gpurun_out/ncu_full2_$WL.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/launches_r01_v2_$WL.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --workload $WL > /dev/null 2>&1
ls -la gpurun_out | tail -5
