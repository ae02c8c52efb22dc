#!/bin/bash
# GaussianCity-scale workloads: ours vs reference (latency-bound regime)
for wl in cfg5_city_16k_540p city_500k_540p cfg2_100k_sh0_512; do
  for impl in reference ours; do
    python bench.py --impl $impl --steps 50 --warmup 5 --no-cpu-baseline --workload $wl 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl', '$impl', 'ms', round(d['ms_per_step'],4), 'fwd', round(d['ms_forward'],4), 'R', d['num_rendered'], 'e2e ms', round(d['e2e']['ms_per_step'],4), {k: round(v,4) for k,v in (d.get('stage_ms') or {}).items()})"
  done
done
