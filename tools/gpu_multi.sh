#!/bin/bash
# run with: gpurun --gpus N -- 'bash tools/gpu_multi.sh N'   (full JSON lines land in gpurun_out/)
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    tools/gpu_shard_check.py cfg3_1M_sh3_1080p 2>&1 | grep -E "^rank" | cut -c1-200
python bench.py --gpus 1 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/scale_n1.json 2>gpurun_out/scale_n1.err
for mode in broadcast replicated; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
      bench.py --gpus $N --steps 10 --warmup 3 --shard-mode $mode > gpurun_out/scale_n${N}_$mode.json 2>gpurun_out/scale_n${N}_$mode.err
done
python - <<PY
import json
for f in ["scale_n1", "scale_n${N}_broadcast", "scale_n${N}_replicated"]:
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "fwd", round(d["ms_forward"], 3))
    except Exception as e:
        print(f, "FAILED", e)
PY
