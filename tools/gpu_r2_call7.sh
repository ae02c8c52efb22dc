#!/bin/bash
# 8 GPUs: scaling bench (peer exchange), both assembly modes are reported by the line itself
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/c7_topo.txt 2>&1
for n in 8 4; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/c7_bench_n$n.err | tail -1 > gpurun_out/c7_bench_n$n.json
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 tools/config5_gstep.py --arm reference > gpurun_out/c7_cfg5_ref.json 2>gpurun_out/c7_cfg5_ref.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 tools/config5_gstep.py --arm ours > gpurun_out/c7_cfg5_ours.json 2>gpurun_out/c7_cfg5_ours.err
