#!/bin/bash
# A/B of blend_bwd v1 vs v2 + tests
timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
for v in v1 v2; do
  GCR_BLEND_BWD=$v python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', round(d['value'],1), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stage_ms'].items()})"
done
GCR_BLEND_BWD=v2 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --workload cfg3_1M_sh3_1080p 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('1M', round(d['value'],1), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['stage_ms'].items()})"
