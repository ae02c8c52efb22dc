#!/bin/bash
# 2 GPUs: validate the select-stage compaction + chunked geometry backward under real NCCL, N=2 bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nccl_stripes.py tests/test_gpu_vs_oracle.py tests/test_gpu_api_edges.py -q -m gpu 2>&1 | tail -15 > gpurun_out/c11_pytest.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/c11_bench_n2.err | tail -1 > gpurun_out/c11_bench_n2.json
