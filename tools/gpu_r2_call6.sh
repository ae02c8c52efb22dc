#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/gpu_nccl_debug.py > gpurun_out/c6_debug.log 2>&1
timeout 900 python -m pytest tests/ -q -m gpu 2>&1 | tail -30 > gpurun_out/c6_pytest.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/c6_bench_n2.err | tail -1 > gpurun_out/c6_bench_n2.json
timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline 2>gpurun_out/c6_bench_n1.err | tail -1 > gpurun_out/c6_bench_n1.json
