"""GPU (round 2): per-call latency of GaussianRasterizerWrapper at GaussianCity's own scale
(16 384 lattice points, K / sensor 960x540) with the reference's sequence of camera
operations (device matmul + device 4x4 inverse + 3 small copies per call) versus
`fast_camera=True` (host math, one packed upload, LRU pose cache) -- and the two images compared.

  gpurun --timeout 600 -- 'python tools/gpu_wrapper_latency.py > gpurun_out/wrapper_latency.log 2>&1'
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import gaussiancity_b200 as g  # noqa: E402
from gaussiancity_b200.synthetic import CITY_K, CITY_SENSOR, city_points  # noqa: E402
from tests.golden.camera.make_wrapper_golden import orbit_poses  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    pts, _, _ = city_points(16384, seed=0, device=dev)
    poses = orbit_poses(n=24)[:24]
    res = {}
    for name, fast in (("reference_sequence", False), ("fast_camera", True)):
        w = g.GaussianRasterizerWrapper(CITY_K, CITY_SENSOR, device=dev, fast_camera=fast)
        with torch.no_grad():
            for pos, quat in poses[:4]:                       # warm-up (module load, allocator)
                img = w(pts, pos, quat)
            torch.cuda.synchronize()
            # (a) a new pose every call
            t0 = time.perf_counter()
            for rep in range(10):
                for pos, quat in poses:
                    img = w(pts, pos + rep * 1e-3, quat)
            torch.cuda.synchronize()
            new_pose_us = (time.perf_counter() - t0) / (10 * len(poses)) * 1e6
            # (b) the same poses again (cache hits on the fast path)
            t0 = time.perf_counter()
            for rep in range(10):
                for pos, quat in poses:
                    img = w(pts, pos, quat)
            torch.cuda.synchronize()
            same_pose_us = (time.perf_counter() - t0) / (10 * len(poses)) * 1e6
            # (c) settings only
            t0 = time.perf_counter()
            for rep in range(10):
                for pos, quat in poses:
                    w._get_gaussian_rasterization_settings(pos + 1.0 + rep * 1e-3, quat)
            torch.cuda.synchronize()
            settings_us = (time.perf_counter() - t0) / (10 * len(poses)) * 1e6
        res[name] = dict(new_pose_us=new_pose_us, same_pose_us=same_pose_us, settings_only_us=settings_us,
                         img=w(pts, *poses[3]).float().cpu().numpy())
        print(name, {k: round(v, 1) for k, v in res[name].items() if k != "img"})
    a, b = res["reference_sequence"]["img"], res["fast_camera"]["img"]
    print("image norm-relative difference:", float(np.linalg.norm(a - b) / max(np.linalg.norm(a), 1e-30)),
          "bit-identical:", bool(np.array_equal(a, b)))


if __name__ == "__main__":
    main()
