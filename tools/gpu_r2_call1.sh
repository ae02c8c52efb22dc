#!/bin/bash
# round-2 call 1: full gpu suite on HEAD, A/B of the prepared variants, host overhead in the 16k regime
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -5 > gpurun_out/c1_pytest.log
bash tools/gpu_ab_r2.sh > gpurun_out/c1_ab.log 2>&1
for wl in cfg5_city_16k_540p cfg4_5M_sh3_1080p; do
  echo "== $wl"; GCR_HOST_TIMING=1 python tools/gpu_overhead.py $wl 2>&1 | tail -25
done > gpurun_out/c1_overhead.log 2>&1
