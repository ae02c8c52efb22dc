#!/bin/bash
# run with: gpurun --gpus N -- 'bash tools/gpu_multi.sh N'
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    tools/gpu_shard_check.py cfg3_1M_sh3_1080p 2>&1 | grep -v Warning | tail -$((N+6))
for n in 1 $N; do
  if [ $n -eq 1 ]; then python bench.py --gpus 1 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline | tee gpurun_out/scale_n1.json | cut -c1-700;
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 \
      bench.py --gpus $n --steps 10 --warmup 3 2>gpurun_out/scale_n$n.err | tee gpurun_out/scale_n$n.json | cut -c1-700; fi
done
