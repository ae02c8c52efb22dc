#!/bin/bash
# round-2 call 3 (2 GPUs): real-NCCL stripes test (peer + collective exchange), N=2 bench, N=1 bench for the emit fix
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/c3_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_nccl_stripes.py -x -q -m gpu 2>&1 | tail -30 > gpurun_out/c3_pytest_nccl.log
for mode in "" "--assemble" "--exchange collective"; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline $mode 2>gpurun_out/c3_bench_n2.err | tail -1 >> gpurun_out/c3_bench_n2.json
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline 2>gpurun_out/c3_bench_n1.err | tail -1 > gpurun_out/c3_bench_n1.json
