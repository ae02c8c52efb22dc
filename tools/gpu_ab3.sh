#!/bin/bash
timeout 300 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -2
for wl in cfg5_city_16k_540p cfg2_100k_sh0_512; do
python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --workload $wl 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl', round(d['ms_per_step'],4), 'fwd', round(d['ms_forward'],4), {k: round(v,4) for k,v in d['stage_ms'].items() if 'sort' in k})"
done
for v in 1024 4096; do
GCR_SORT_TILE=$v python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('5M tile $v', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stage_ms'].items() if 'sort' in k})"
done
