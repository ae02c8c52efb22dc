#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/gpu_peer_diag.py > gpurun_out/c4_peer_diag.log 2>&1
timeout 600 python -m pytest tests/test_gpu_nccl_stripes.py tests/test_gpu_vs_reference.py tests/test_gpu_vs_oracle.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/c4_pytest.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/c4_bench_n2.err | tail -1 > gpurun_out/c4_bench_n2.json
timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline 2>gpurun_out/c4_bench_n1.err | tail -1 > gpurun_out/c4_bench_n1.json
