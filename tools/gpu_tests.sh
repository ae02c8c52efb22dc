#!/bin/bash
# Run on the B200 box through gpurun: GPU parity tests; logs to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv | tee gpurun_out/gpu.txt
nproc | tee -a gpurun_out/gpu.txt
timeout 1500 python -m pytest tests/ -x -q -m gpu -s 2>&1 | tee gpurun_out/pytest_gpu.log | tail -80
