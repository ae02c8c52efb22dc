#!/bin/bash
# 8 GPUs: scaling bench N=8 and N=4 (peer exchange, balanced stripes), config 5 under DDP (3 arms, twice, alternating)
mkdir -p gpurun_out
for n in 8 4; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/c12_bench_n$n.err | tail -1 > gpurun_out/c12_bench_n$n.json
done
for rep in 1 2; do
for arm in reference ours ours_wrapper; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2952$rep tools/config5_gstep.py --arm $arm --steps 100 --warmup 10 > gpurun_out/c12_cfg5_${arm}_$rep.json 2>gpurun_out/c12_cfg5_${arm}_$rep.err
done
done
