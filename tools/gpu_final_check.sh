#!/bin/bash
# Last call of round 2: the full GPU suite on the final build (PDL on, grid encoder in), smoke(), the
# headline bench on both arms, the small-frame lines, the grid-encoder leg, fresh golden vectors for the
# grid encoder (with the device's level scales) and the ncu launch list of the headline command.
mkdir -p gpurun_out
O=gpurun_out/fin
timeout 600 python -m pytest tests/ -q -m gpu 2>&1 | tail -15 > ${O}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1
timeout 200 python bench.py --impl reference --steps 20 --warmup 5 2>${O}_bench_reference.err | tail -1 > ${O}_bench_reference_cfg4_5M_sh3_1080p.json
timeout 300 python bench.py --steps 20 --warmup 5 2>${O}_bench_ours.err | tail -1 > ${O}_bench_ours_cfg4_5M_sh3_1080p.json
for impl in reference ours; do
  timeout 120 python bench.py --impl $impl --workload cfg5_city_16k_540p --steps 300 --warmup 30 --no-cpu-baseline 2>>${O}_bench_small.err | tail -1 > ${O}_bench_${impl}_cfg5_city_16k_540p.json
  timeout 120 python bench.py --impl $impl --workload grid_encoder --steps 50 --warmup 5 2>>${O}_bench_grid.err | tail -1 > ${O}_bench_grid_${impl}.json
done
timeout 100 python tests/golden/make_golden_grid.py gpurun_out/golden_grid2 > ${O}_golden.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/r02_launches_final.csv \
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > ${O}_ncu_launch.log 2>&1
