"""Generates tests/golden/camera/wrapper_settings.npz from the UNMODIFIED reference camera adapter
(`GaussianRasterizerWrapper`, /root/reference/extensions/diff_gaussian_rasterization/__init__.py:
276-402) run on CPU tensors in this container: for an orbit of poses around the GoogleEarth-style
camera (K / sensor 960x540 from gaussiancity_b200.synthetic) it records the settings the reference
hands to the rasterizer -- view_matrix, proj_matrix, campos, tan(fov/2) -- plus the projection P.

Needs /root/reference and oracle/_ref (the reference module imports its native extension at import
time); the committed .npz is what the CPU tests read.  Run:  python tests/golden/camera/make_wrapper_golden.py
"""
import importlib.util
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, ROOT)
REF_PKG = "/root/reference/extensions/diff_gaussian_rasterization/__init__.py"


def look_at_quat_xyzw(cam_pos, target):
    """(qx,qy,qz,qw) of the rotation whose columns are [F|R|U] (the convention of
    scripts/dataset_generator.py:1071-1085, restated with the repo's own helper)."""
    from gaussiancity_b200.synthetic import _matrix_to_quat_xyzw
    fwd = np.asarray(target, np.float64) - np.asarray(cam_pos, np.float64)
    fwd /= np.linalg.norm(fwd)
    right = np.cross(np.array([0.0, 0.0, 1.0]), fwd)
    right /= np.linalg.norm(right)
    up = np.cross(fwd, right)
    return _matrix_to_quat_xyzw(np.stack([fwd, right, up], axis=1))


def orbit_poses(n=8, radius=520.0, altitude=300.0, centre=(0.0, 0.0, 1.0)):
    poses = []
    for i in range(n):
        th = 2 * math.pi / n * i + 0.1
        pos = np.array([centre[0] + radius * math.cos(th), centre[1] + radius * math.sin(th), altitude])
        poses.append((pos, look_at_quat_xyzw(pos, centre)))
    # one float32 pose and one un-normalised quaternion, as callers may pass either
    poses.append((poses[1][0].astype(np.float32), poses[1][1].astype(np.float32)))
    poses.append((poses[2][0], poses[2][1] * 3.0))
    return poses


def main():
    from gaussiancity_b200.synthetic import CITY_K, CITY_SENSOR
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    spec = importlib.util.spec_from_file_location("dgr_reference", REF_PKG)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    w = ref.GaussianRasterizerWrapper(CITY_K, CITY_SENSOR, device=torch.device("cpu"))
    poses = orbit_poses()
    out = dict(K=np.asarray(CITY_K, np.float64), sensor=np.asarray(CITY_SENSOR, np.int64),
               P=w.P.numpy(), cam_pos=[], cam_quat=[], view=[], proj=[], campos=[])
    for pos, quat in poses:
        st = w._get_gaussian_rasterization_settings(pos, quat)
        out["cam_pos"].append(np.asarray(pos, np.float64))
        out["cam_quat"].append(np.asarray(quat, np.float64))
        out["view"].append(st.view_matrix.numpy().copy())
        out["proj"].append(st.proj_matrix.numpy().copy())
        out["campos"].append(st.campos.numpy().copy())
    out["tanfov"] = np.array([st.tanfovx, st.tanfovy], np.float64)
    out["pose_is_f32"] = np.array([False] * 8 + [True, False])
    for k in ("cam_pos", "cam_quat", "view", "proj", "campos"):
        out[k] = np.stack(out[k])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "wrapper_settings.npz")
    np.savez(path, **out)
    print("wrote", path, {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
