#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nccl_stripes.py tests/test_gpu_vs_oracle.py -q -m gpu 2>&1 | tail -8 > gpurun_out/c17_pytest.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/c17_bench_n2.err | tail -1 > gpurun_out/c17_bench_n2.json
timeout 200 python tools/gpu_ncu_striped.py --world 8 --rank 3 --steps 5 > gpurun_out/c17_striped_rank3of8.log 2>&1
