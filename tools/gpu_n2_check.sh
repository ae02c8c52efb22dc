#!/bin/bash
# 2-GPU sanity of the final build: the real-NCCL stripes test and the striped bench with its bit-identity gate.
mkdir -p gpurun_out
O=gpurun_out/n2
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
  bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>${O}_bench.err | tail -1 > ${O}_bench.json
timeout 100 python -m pytest tests/test_gpu_nccl_stripes.py -q -m gpu 2>&1 | tail -4 > ${O}_pytest.log
