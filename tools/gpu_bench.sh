#!/bin/bash
# First measurement pass on the B200 box: bench lines (ours + reference), ncu launch list and a
# full capture of the blend kernels.
mkdir -p gpurun_out
WL=${1:-cfg3_1M_sh3_1080p}
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --impl reference --steps 5 --warmup 3 --workload $WL 2> gpurun_out/bench_ref_$WL.err | tee gpurun_out/bench_ref_$WL.json
python bench.py --steps 5 --warmup 3 --workload $WL 2> gpurun_out/bench_ours_$WL.err | tee gpurun_out/bench_ours_$WL.json
tail -5 gpurun_out/bench_ours_$WL.err
