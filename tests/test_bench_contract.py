"""CPU: pieces of the bench.py contract that can be checked without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_refuses_to_run_without_cuda():
    """No CPU path: on a box without a CUDA device bench.py reports an error instead of timing
    anything (the product must fail loudly, never fall back)."""
    import torch
    if torch.cuda.is_available():
        return
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--workload", "tiny"],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode != 0
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert "no CUDA device" in line["error"]


def test_default_workload_is_the_headline_config():
    sys.path.insert(0, ROOT)
    import bench
    P, W, H, deg, use_sh = bench.WORKLOADS[bench.DEFAULT_WORKLOAD]
    assert (P, W, H, deg, use_sh) == (5_000_000, 1920, 1080, 3, True)   # BASELINE.json configs[3]
    assert set(bench.WORKLOADS) >= {"cfg2_100k_sh0_512", "cfg3_1M_sh3_1080p", "cfg5_city_16k_540p"}


def test_traffic_file_matches_default_workload():
    d = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
    sys.path.insert(0, ROOT)
    import bench
    assert d["workload"] == bench.DEFAULT_WORKLOAD
    for k in ("blend_fwd", "blend_bwd"):
        assert d["kernels"][k]["dram_bytes"] > 1e8


def test_grid_encoder_leg_refuses_to_run_without_cuda():
    """bench.py --workload grid_encoder has no CPU path either."""
    import torch
    if torch.cuda.is_available():
        return
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--workload", "grid_encoder"],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode != 0
    assert "no CUDA device" in json.loads(p.stdout.strip().splitlines()[-1])["error"]
