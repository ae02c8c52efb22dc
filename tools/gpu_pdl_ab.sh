#!/bin/bash
# A/B of programmatic dependent launch (GCR_PDL=0/1) + the grid encoder's remaining GPU tests and its
# fused-backward arm.  Everything lands in gpurun_out/pdl_*.
mkdir -p gpurun_out
O=gpurun_out/pdl
timeout 400 python -m pytest tests/test_gpu_grid_encoder.py "tests/test_gpu_vs_reference.py::test_other_launch_mode_matches_reference" \
  "tests/test_gpu_api_edges.py::test_launch_modes_agree_on_small_frames" -q -m gpu 2>&1 | tail -40 > ${O}_pytest.log
for pdl in 0 1 0 1; do
  GCR_PDL=$pdl timeout 120 python bench.py --workload cfg5_city_16k_540p --steps 300 --warmup 30 --no-cpu-baseline 2>>${O}_bench.err | tail -1 >> ${O}_bench_16k_pdl${pdl}.json
done
for pdl in 0 1; do
  GCR_PDL=$pdl timeout 120 python bench.py --workload cfg2_100k_sh0_512 --steps 100 --warmup 10 --no-cpu-baseline --no-e2e 2>>${O}_bench.err | tail -1 > ${O}_bench_100k_pdl${pdl}.json
  GCR_PDL=$pdl timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>>${O}_bench.err | tail -1 > ${O}_bench_5M_pdl${pdl}.json
done
timeout 120 python bench_grid_encoder.py --impl ours --fused --steps 50 --warmup 5 --no-cpu-baseline 2>${O}_grid_fused.err | tail -2 > ${O}_grid_fused.json
