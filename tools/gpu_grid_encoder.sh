#!/bin/bash
# One GPU call for the hash-grid encoder: golden vectors from the unmodified reference
# extension, the GPU parity tests, bench lines (ours / reference), an ncu metrics pass over both arms'
# kernels and the config-5 harness with POS_EMD=HASH_GRID.  Everything lands in gpurun_out/grid_*.
mkdir -p gpurun_out
O=gpurun_out/grid
timeout 120 python tests/golden/make_golden_grid.py gpurun_out/golden_grid > ${O}_golden.log 2>&1
timeout 300 python -m pytest tests/test_gpu_grid_encoder.py -q -m gpu -x 2>&1 | tail -40 > ${O}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1
for impl in ours reference; do
  timeout 200 python bench.py --workload grid_encoder --impl $impl --steps 50 --warmup 5 2>${O}_bench_${impl}.err | tail -1 > ${O}_bench_${impl}.json
done
timeout 200 python bench_grid_encoder.py --impl both --points 262144 --steps 20 --warmup 3 --no-cpu-baseline 2>${O}_bench_256k.err | tail -2 > ${O}_bench_256k.json
timeout 200 python bench_grid_encoder.py --impl both --dims 3 --steps 50 --warmup 5 --no-cpu-baseline 2>${O}_bench_d3.err | tail -2 > ${O}_bench_d3.json
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,launch__registers_per_thread,smsp__inst_executed.sum \
  --clock-control none -k regex:'grid|kernel_' --csv --log-file ${O}_ncu_metrics.csv \
  python bench_grid_encoder.py --impl both --once > ${O}_ncu.log 2>&1
for arm in reference ours; do
  timeout 200 python tools/config5_gstep.py --arm $arm --pos-emd HASH_GRID --encoder global --steps 100 --warmup 10 > ${O}_cfg5_hash_${arm}.json 2>${O}_cfg5_hash_${arm}.err
done
