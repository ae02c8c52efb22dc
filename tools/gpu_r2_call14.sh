#!/bin/bash
mkdir -p gpurun_out
timeout 250 compute-sanitizer --tool racecheck --print-limit 4 python tools/gpu_sanitize.py 2>&1 | grep -E "RACECHECK SUMMARY|Race reported|and (Read|Write)|^ok" | cut -c1-220 | sort | uniq -c | sort -rn | head -12 > gpurun_out/c14_racecheck.log
timeout 900 python -m pytest tests/ -q -m gpu 2>&1 | tail -25 > gpurun_out/c14_pytest.log
for v in prefetch noprefetch; do
  if [ $v = noprefetch ]; then export GCR_NO_PREFETCH=1; else unset GCR_NO_PREFETCH; fi
  for wl in cfg4_5M_sh3_1080p cfg3_1M_sh3_1080p; do
    timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --workload $wl 2>gpurun_out/c14_bench_${v}_$wl.err | tail -1 > gpurun_out/c14_bench_${v}_$wl.json
  done
done
unset GCR_NO_PREFETCH
timeout 200 python tools/gpu_ncu_striped.py --world 8 --rank 3 > gpurun_out/c14_striped_rank3of8.log 2>&1
timeout 200 python tools/gpu_ncu_striped.py --world 8 --rank 0 > gpurun_out/c14_striped_rank0of8.log 2>&1
