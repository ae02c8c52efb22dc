#!/bin/bash
# Round-2 final multi-GPU record on an N-GPU box (N = $1): scaling bench at N (both shard modes, both
# exchanges) and, for N = 8, the config-5 harness under DDP.
N=${1:-8}
mkdir -p gpurun_out
O=gpurun_out/final
run() { # tag, extra flags
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline $2 2>${O}_scale_n${N}_$1.err | tail -1 > ${O}_scale_n${N}_$1.json
}
run peer ""
run peer_assemble "--assemble --no-e2e"
run collective "--exchange collective --no-e2e"
run broadcast "--shard-mode broadcast --no-e2e"
if [ "$N" = "8" ]; then
  for arm in reference ours ours_wrapper ours_fused; do
    timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/config5_gstep.py --arm $arm --steps 200 --warmup 20 > ${O}_cfg5_${arm}_n8.json 2>${O}_cfg5_${arm}_n8.err
  done
fi
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_gpu_nccl_stripes.py -q -m gpu 2>&1 | tail -4 > ${O}_pytest_nccl.log
fi
