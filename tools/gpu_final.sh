#!/bin/bash
# Round-end evidence run on one B200: GPU tests, both bench arms, ncu launch list + full capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader | tee gpurun_out/gpu.txt; nproc | tee -a gpurun_out/gpu.txt
timeout 600 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_tail.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-300
python bench.py --impl reference > gpurun_out/BENCH_ref.json 2> gpurun_out/BENCH_ref.err
python bench.py > gpurun_out/BENCH_ours.json 2> gpurun_out/BENCH_ours.err
python - <<PY
import json
for f in ("BENCH_ref", "BENCH_ours"):
    d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    print(f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "fwd", round(d["ms_forward"], 3),
          "e2e", round(d["e2e"]["value"], 1), round(d["e2e"]["ms_per_step"], 2), "clocks", d.get("clocks"))
    if "stage_ms" in d: print("   ", {k: round(v, 3) for k, v in d["stage_ms"].items()}, d["roofline"]["frac"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 300 --csv --log-file gpurun_out/launches_r01_v3.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"blend_|preprocess_|onesweep|ranges_gather" -s 33 -c 11 -f -o gpurun_out/prof_r01_v3 \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full3.log 2>&1
ls -la gpurun_out | grep -E "v3|BENCH"
