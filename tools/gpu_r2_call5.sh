#!/bin/bash
mkdir -p gpurun_out
timeout 200 compute-sanitizer --tool memcheck --print-limit 8 python tools/gpu_sanitize.py > gpurun_out/c5_memcheck.log 2>&1
timeout 900 python -m pytest tests/ -q -m gpu 2>&1 | tail -40 > gpurun_out/c5_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/gpu_peer_diag.py > gpurun_out/c5_peer_diag.log 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/c5_bench_n2.err | tail -1 > gpurun_out/c5_bench_n2.json
for wl in cfg4_5M_sh3_1080p cfg5_city_16k_540p; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --workload $wl 2>gpurun_out/c5_bench_n1_$wl.err | tail -1 > gpurun_out/c5_bench_n1_$wl.json
done
