#!/bin/bash
# Round-2 A/B of the experimental variants prepared (compile-checked only) at the end of round 1:
#   GCR_SORT_ORDER=counts         early-counts onesweep pass (binning.cu)
#   GCR_SORT_SPIN_NS=<ns>         __nanosleep back-off in the look-back spin
#   GCR_BWD_MATH=approx           ex2.approx / rcp.approx in blend_bwd_kernel_v2 (blend_bwd.cu)
#   GCR_BLEND_FWD=trim            hoisted address/position arithmetic in blend_fwd_kernel
# For each variant: the bit-exactness suite against the reference extension (sort order, keys,
# gradients <= 1e-4), then the headline bench with per-stage times.  One gpurun call:
#   gpurun --timeout 1500 -- 'bash tools/gpu_ab_r2.sh > gpurun_out/ab_r2.log 2>&1'
run() {
  local tag="$1"; shift
  echo "=== $tag"
  env "$@" timeout 600 python -m pytest tests/test_gpu_vs_reference.py tests/test_gpu_vs_oracle.py -x -q -m gpu 2>&1 | tail -2
  for wl in cfg4_5M_sh3_1080p cfg3_1M_sh3_1080p cfg5_city_16k_540p; do
    env "$@" python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --workload $wl 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$tag', '$wl', round(d['value'],1), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stage_ms'].items()})"
  done
}
run baseline            GCR_AB=0
run sort_counts         GCR_SORT_ORDER=counts
run sort_counts_spin100 GCR_SORT_ORDER=counts GCR_SORT_SPIN_NS=100
run bwd_approx          GCR_BWD_MATH=approx
run fwd_trim            GCR_BLEND_FWD=trim
run all                 GCR_SORT_ORDER=counts GCR_BWD_MATH=approx GCR_BLEND_FWD=trim
