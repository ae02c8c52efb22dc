"""CPU: host-side mirror of the reference operator surface (DGR/__init__.py) -- argument
validation, settings tuple, camera adapter math, refusal to run without CUDA."""
import math

import numpy as np
import pytest
import torch

import gaussiancity_b200 as g
from gaussiancity_b200.synthetic import CITY_K, CITY_SENSOR, city_points, uniform_scene


def test_settings_fields_match_reference_order():
    assert g.GaussianRasterizationSettings._fields == (
        "img_h", "img_w", "tanfovx", "tanfovy", "bg", "scale_modifier", "view_matrix", "proj_matrix",
        "sh_degree", "campos", "prefiltered", "debug")


def _settings(s):
    return g.GaussianRasterizationSettings(s.img_h, s.img_w, s.tanfovx, s.tanfovy, s.bg, 1.0,
                                           s.view_matrix, s.proj_matrix, s.sh_degree, s.campos, False, False)


def test_rasterizer_argument_validation():
    s = uniform_scene(10, 32, 32, seed=0)
    r = g.GaussianRasterizer(_settings(s))
    m2 = torch.zeros_like(s.means3D)
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(s.means3D, m2, s.opacities, shs=None, colors_precomp=None, scales=s.scales, rotations=s.rotations)
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(s.means3D, m2, s.opacities, shs=s.shs, colors_precomp=torch.zeros(10, 3), scales=s.scales,
          rotations=s.rotations)
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(s.means3D, m2, s.opacities, shs=s.shs, scales=s.scales, rotations=None)
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(s.means3D, m2, s.opacities, shs=s.shs, scales=s.scales, rotations=s.rotations,
          cov3D_precomp=torch.zeros(10, 6))


def test_cpu_tensors_are_refused_not_silently_computed(built_lib):
    s = uniform_scene(10, 32, 32, seed=0)
    r = g.GaussianRasterizer(_settings(s))
    with pytest.raises(RuntimeError, match="no CPU path"):
        r(s.means3D, torch.zeros_like(s.means3D), s.opacities, shs=s.shs, scales=s.scales,
          rotations=s.rotations)
    with pytest.raises(RuntimeError, match="no CPU path"):
        g.mark_visible(s.means3D, s.view_matrix, s.proj_matrix)


def test_bad_means_shape_raises(built_lib):
    from gaussiancity_b200 import ext
    e = torch.Tensor([])
    with pytest.raises(RuntimeError, match=r"means3D must have dimensions \(num_points, 3\)"):
        ext.rasterize_gaussians(torch.zeros(3), torch.zeros(5, 4), e, torch.zeros(5, 1), e, e, 1.0, e,
                                torch.eye(4), torch.eye(4), 1.0, 1.0, 8, 8, e, 0, torch.zeros(3), False, False)


def test_wrapper_camera_matches_reference_formulas():
    """fov / projection / w2c as DGR/__init__.py:326-402 computes them (restated here)."""
    w = g.GaussianRasterizerWrapper(CITY_K, CITY_SENSOR, device=torch.device("cpu"))
    fx, fy, cx, cy = CITY_K[0, 0], CITY_K[1, 1], CITY_K[0, 2], CITY_K[1, 2]
    assert w.fov_x == pytest.approx(2 * math.atan2(960, 2 * fx))
    assert w.fov_y == pytest.approx(2 * math.atan2(540, 2 * fy))
    P = w.P.numpy()
    zn, zf = 0.01, 50000.0
    exp = np.zeros((4, 4), np.float32)
    exp[0, 0], exp[1, 1] = 2 * fx / 960, 2 * fy / 540
    exp[0, 2], exp[1, 2] = 2 * cx / 960 - 1, 2 * cy / 540 - 1
    exp[2, 2], exp[3, 2], exp[2, 3] = -(zf + zn) / (zf - zn), -1.0, -2 * zf * zn / (zf - zn)
    assert np.array_equal(P, exp)
    _, cam_pos, cam_quat = city_points(10, seed=1)
    st = w._get_gaussian_rasterization_settings(cam_pos, cam_quat)
    assert (st.img_w, st.img_h) == CITY_SENSOR and st.sh_degree == 0 and st.scale_modifier == 1.0
    w2c = st.view_matrix.T.numpy().astype(np.float64)
    R = w2c[:3, :3]
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-6)          # rigid
    assert np.allclose(w2c[:3, :3] @ cam_pos + w2c[:3, 3], 0, atol=1e-3)  # camera centre -> origin
    assert np.allclose(st.campos.numpy(), cam_pos, atol=1e-2)
    assert torch.allclose(st.proj_matrix, st.view_matrix @ w.P.T)
    # the scene centre is in front of the camera (z_view > 0.2) and has NEGATIVE clip w
    p = np.array([0.0, 0.0, 10.0, 1.0])
    assert (w2c @ p)[2] > 0.2
    assert (st.proj_matrix.T.numpy().astype(np.float64) @ p)[3] < 0


def test_wrapper_requires_14_channels():
    w = g.GaussianRasterizerWrapper(CITY_K, CITY_SENSOR, device=torch.device("cpu"))
    with pytest.raises(AssertionError, match="14 channels"):
        w(torch.zeros(5, 13), np.zeros(3), np.array([0, 0, 0, 1.0]))


def test_compat_module_exports_reference_names(built_lib):
    import importlib
    import os
    import sys
    d = os.path.join(os.path.dirname(g.__file__), "compat")
    sys.path.insert(0, d)
    try:
        m = importlib.import_module("diff_gaussian_rasterization_ext")
        for n in ("rasterize_gaussians", "rasterize_gaussians_backward", "mark_visible"):
            assert callable(getattr(m, n))
    finally:
        sys.path.remove(d)
        sys.modules.pop("diff_gaussian_rasterization_ext", None)


def test_bench_kernel_count_formula():
    import bench
    # 1080p: 8160 tiles -> 13 bits -> 2 tile passes; 1 + 5 + 1 + 2 + 1 + 1 (+2 backward)
    assert bench.tile_sort_passes(1920, 1080) == 2
    assert bench.kernels_per_step(1920, 1080) == 13
    assert bench.tile_sort_passes(128, 128) == 1


# ---- GaussianRasterizerWrapper camera path against the reference's own settings ----

def _wrapper_golden():
    import os
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "camera", "wrapper_settings.npz"))


def _golden_pose(gold, i):
    pos, quat = gold["cam_pos"][i], gold["cam_quat"][i]
    if gold["pose_is_f32"][i]:
        pos, quat = pos.astype(np.float32), quat.astype(np.float32)
    return pos, quat


@pytest.mark.parametrize("fast", [False, True], ids=["reference_sequence", "fast_camera"])
def test_wrapper_settings_match_reference_golden(fast):
    """tests/golden/camera/wrapper_settings.npz holds what the UNMODIFIED reference wrapper hands to the
    rasterizer for an orbit of poses (generated by tests/golden/camera/make_wrapper_golden.py)."""
    gold = _wrapper_golden()
    w = g.GaussianRasterizerWrapper(gold["K"], tuple(int(v) for v in gold["sensor"]),
                                    device=torch.device("cpu"), fast_camera=fast)
    assert np.array_equal(w.P.numpy(), gold["P"])
    for i in range(gold["cam_pos"].shape[0]):
        pos, quat = _golden_pose(gold, i)
        st = w._get_gaussian_rasterization_settings(pos, quat)
        assert (st.tanfovx, st.tanfovy) == (gold["tanfov"][0], gold["tanfov"][1])
        # host expression is the reference's own: bit-identical view matrix on both paths
        assert np.array_equal(st.view_matrix.numpy(), gold["view"][i])
        # proj: a 4-term fp32 dot product per element -- within 1 ulp of the row's magnitude
        tol = 2.0 ** -22 * np.abs(gold["proj"][i]).max(axis=1, keepdims=True)
        assert (np.abs(st.proj_matrix.numpy().astype(np.float64) - gold["proj"][i]) <= tol).all()
        # campos: the reference inverts the fp32 view matrix; the true centre is cam_pos
        assert np.allclose(st.campos.numpy(), gold["campos"][i], atol=2e-3)
        assert np.allclose(st.campos.numpy(), gold["cam_pos"][i], atol=2e-3)
        assert st.bg.shape == (3,) and float(st.bg.abs().sum()) == 0.0
        for t in (st.view_matrix, st.proj_matrix, st.campos, st.bg):
            assert t.dtype == torch.float32
            assert t.is_contiguous() or not fast   # (the reference hands over a transposed view)
        if fast:
            assert np.array_equal(st.campos.numpy(), pos.astype(np.float32))


def test_fast_camera_cache_uploads_once_per_pose():
    gold = _wrapper_golden()
    w = g.GaussianRasterizerWrapper(gold["K"], (960, 540), device=torch.device("cpu"), fast_camera=True)
    c = w._pose_cache
    poses = [_golden_pose(gold, i) for i in range(4)]
    first = [w._get_gaussian_rasterization_settings(*p) for p in poses]
    assert c.uploads == 4
    again = [w._get_gaussian_rasterization_settings(*p) for p in poses]
    assert c.uploads == 4                                     # all hits
    for a, b in zip(first, again):
        assert a.view_matrix.data_ptr() == b.view_matrix.data_ptr()
    # torch tensors as pose, same values -> same entry as the numpy pose
    w._get_gaussian_rasterization_settings(torch.from_numpy(poses[0][0]), torch.from_numpy(poses[0][1]))
    assert c.uploads == 4
    # a different dtype of the same pose is a different key (the reference computes in that dtype)
    w._get_gaussian_rasterization_settings(poses[0][0].astype(np.float32), poses[0][1])
    assert c.uploads == 5
    # LRU eviction
    c.capacity = 2
    for i in range(4, 8):
        w._get_gaussian_rasterization_settings(*_golden_pose(gold, i))
    assert len(c._lru) == 2
    # 16-byte alignment of every section of the packed upload
    st = w._get_gaussian_rasterization_settings(*poses[1])
    base = st.view_matrix.data_ptr()
    assert (st.proj_matrix.data_ptr() - base, st.campos.data_ptr() - base, st.bg.data_ptr() - base) == (64, 128, 144)


def test_fast_camera_renders_the_same_image_as_the_reference_sequence():
    """End to end through the CPU oracle (no GPU here): the settings of both camera paths give the
    same frame within the colour bar (<= 1e-4; the matrices differ by at most the last bit)."""
    from gaussiancity_b200.synthetic import city_points
    from oracle import oracle
    pts, cam_pos, cam_quat = city_points(600, seed=3)
    imgs = []
    for fast in (False, True):
        w = g.GaussianRasterizerWrapper(CITY_K, CITY_SENSOR, device=torch.device("cpu"), fast_camera=fast)
        st = w._get_gaussian_rasterization_settings(cam_pos, cam_quat)
        p = pts.numpy()
        r = oracle.forward(p[:, 0:3], p[:, 3:4], p[:, 4:7], p[:, 7:11], st.view_matrix.contiguous().numpy(),
                           st.proj_matrix.contiguous().numpy(), st.campos.numpy(), st.img_w, st.img_h,
                           st.tanfovx, st.tanfovy, st.bg.numpy(), colors_precomp=p[:, 11:14])
        assert r.num_rendered > 0
        imgs.append(np.asarray(r.color, np.float64))
    den = np.linalg.norm(imgs[0])
    assert den > 0 and np.linalg.norm(imgs[0] - imgs[1]) / den <= 1e-4


# ---- adapter fusion: crop of the flipped output frame -> window of the unflipped render ----

def test_crop_to_window_commutes_with_the_wrapper_flips():
    from gaussiancity_b200.adapter import crop_to_window
    W, H = 960, 540
    img = np.arange(W * H).reshape(H, W)
    for flip_lr in (False, True):
        for flip_ud in (False, True):
            out = img[:, ::-1] if flip_lr else img
            out = out[::-1, :] if flip_ud else out            # what the wrapper returns
            for crop in (dict(x=160, y=46, w=640, h=448), dict(x=0, y=0, w=1, h=1), dict(x=959, y=539, w=1, h=1),
                         dict(x=7, y=3, w=333, h=210), None):
                x, y, w, h = crop_to_window(W, H, crop, flip_lr, flip_ud)
                win = img[y:y + h, x:x + w]                   # what the rasterizer renders
                win = win[:, ::-1] if flip_lr else win
                win = win[::-1, :] if flip_ud else win
                c = crop or dict(x=0, y=0, w=W, h=H)
                assert np.array_equal(win, out[c["y"]:c["y"] + c["h"], c["x"]:c["x"] + c["w"]])
    for bad in (dict(x=-1, y=0, w=10, h=10), dict(x=0, y=0, w=961, h=10), dict(x=0, y=100, w=10, h=441), dict(x=0, y=0, w=0, h=5)):
        with pytest.raises(ValueError):
            crop_to_window(W, H, bad, True, False)


def test_package_reports_its_host_binding():
    import gaussiancity_b200 as g
    assert g.HOST_BINDING in ("native", "ctypes")
    for n in ("rasterize_gaussians", "rasterize_gaussians_backward", "mark_visible"):
        assert callable(getattr(g.dgr_ext, n))
