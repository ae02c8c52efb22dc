#!/bin/bash
timeout 300 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -2
GCR_SORT_TILE=8192 timeout 300 python -m pytest tests/test_gpu_vs_reference.py -x -q -m gpu 2>&1 | tail -1
for v in 4096 8192; do
  GCR_SORT_TILE=$v python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('tile $v', round(d['value'],1), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stage_ms'].items() if 'sort' in k})"
  GCR_SORT_TILE=$v python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --workload cfg5_city_16k_540p 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  cfg5 tile $v', round(d['ms_per_step'],4), {k: round(v,4) for k,v in d['stage_ms'].items() if 'sort' in k})"
done
