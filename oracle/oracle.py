"""ctypes front-end of the CPU oracle (oracle/gs_oracle.c).  TEST INFRASTRUCTURE -- see the
header of gs_oracle.c.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; gaussiancity_b200 never does.

Two builds of the same source: fp32 (`-ffp-contract=off`, every op rounds to float: the plain
restatement of the reference's fp32 math) and fp64 (arbiter for gradient noise).
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "gs_oracle.c")
BUILD = os.path.join(HERE, "_build")


def build(force=False):
    os.makedirs(BUILD, exist_ok=True)
    out = {}
    for name, dbl in (("f32", 0), ("f64", 1)):
        so = os.path.join(BUILD, f"libgs_oracle_{name}.so")
        if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(SRC):
            cmd = ["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-fno-fast-math",
                   f"-DGSO_DOUBLE={dbl}", SRC, "-o", so, "-lm"]
            subprocess.check_call(cmd)
        out[name] = so
    return out


_libs = {}


def _lib(precision):
    if precision not in _libs:
        so = build()[precision]
        l = ctypes.CDLL(so)
        l.gso_preprocess.restype = ctypes.c_long
        l.gso_bin.restype = ctypes.c_int
        l.gso_render.restype = None
        l.gso_render_backward.restype = None
        l.gso_geometry_backward.restype = None
        _libs[precision] = l
    return _libs[precision]


def _f32(a):
    return None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class OracleResult:
    pass


def set_smooth(on, precision="f64"):
    """Test-only: drop the alpha/T thresholds and widen the bounding square (see gs_oracle.c)."""
    _lib(precision).gso_set_smooth(1 if on else 0)


def forward(means3D, opacities, scales, rotations, view_matrix, proj_matrix, campos, img_w, img_h,
            tanfovx, tanfovy, bg, shs=None, colors_precomp=None, sh_degree=0, scale_modifier=1.0,
            cov3D_precomp=None, precision="f32"):
    """Full forward. All inputs numpy/array-like fp32 in the reference's layouts
    (matrices transposed as GaussianRasterizationSettings carries them). Returns an OracleResult
    with color[3,H,W], radii, num_rendered and every intermediate (state for backward)."""
    l = _lib(precision)
    real = np.float32 if precision == "f32" else np.float64
    means3D = _f32(means3D); P = means3D.shape[0]
    opacities = _f32(opacities).reshape(-1)
    scales, rotations = _f32(scales), _f32(rotations)
    shs, colors_precomp, cov3D_precomp = _f32(shs), _f32(colors_precomp), _f32(cov3D_precomp)
    V, PM = _f32(view_matrix).reshape(-1), _f32(proj_matrix).reshape(-1)
    campos, bg = _f32(campos).reshape(-1), _f32(bg).reshape(-1)
    M = 0 if shs is None else shs.shape[1]
    W, H = int(img_w), int(img_h)
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    r = OracleResult()
    r.P, r.W, r.H, r.M, r.D, r.precision = P, W, H, M, int(sh_degree), precision
    r.inputs = dict(means3D=means3D, opacities=opacities, scales=scales, rotations=rotations, shs=shs,
                    colors_precomp=colors_precomp, cov3D_precomp=cov3D_precomp, V=V, PM=PM,
                    campos=campos, bg=bg, tanfovx=float(tanfovx), tanfovy=float(tanfovy),
                    scale_modifier=float(scale_modifier))
    r.radii = np.zeros(P, np.int32)
    r.depths = np.zeros(P, real); r.means2D = np.zeros((P, 2), real)
    r.cov3D = np.zeros((P, 6), real); r.conic_opacity = np.zeros((P, 4), real)
    r.rgb = np.zeros((P, 3), real); r.clamped = np.zeros((P, 3), np.uint8)
    r.tiles_touched = np.zeros(P, np.uint32)
    R = l.gso_preprocess(P, int(sh_degree), M, W, H, _p(means3D), _p(shs), _p(colors_precomp),
                         _p(opacities), _p(scales), ctypes.c_float(scale_modifier), _p(rotations),
                         _p(cov3D_precomp), _p(V), _p(PM), _p(campos), ctypes.c_float(tanfovx),
                         ctypes.c_float(tanfovy), _p(r.radii), _p(r.depths), _p(r.means2D),
                         _p(r.cov3D), _p(r.conic_opacity), _p(r.rgb), _p(r.clamped),
                         _p(r.tiles_touched))
    r.num_rendered = int(R)
    r.keys = np.zeros(max(R, 1), np.uint64)[:R]
    r.point_list = np.zeros(max(R, 1), np.uint32)[:R]
    r.ranges = np.zeros((tiles, 2), np.uint32)
    keys_buf = np.zeros(max(R, 1), np.uint64); pl_buf = np.zeros(max(R, 1), np.uint32)
    rc = l.gso_bin(P, W, H, _p(r.radii), _p(r.depths), _p(r.means2D), _p(r.tiles_touched),
                   ctypes.c_long(R), _p(keys_buf), _p(pl_buf), _p(r.ranges))
    if rc != 0:
        raise RuntimeError(f"oracle binning failed ({rc})")
    r.keys, r.point_list = keys_buf[:R], pl_buf[:R]
    r.colors = r.rgb if colors_precomp is None else colors_precomp.astype(real)
    r.final_T = np.zeros((H, W), real); r.n_contrib = np.zeros((H, W), np.uint32)
    r.color = np.zeros((3, H, W), real)
    l.gso_render(W, H, _p(r.ranges), _p(pl_buf), _p(r.means2D), _p(r.colors), _p(r.conic_opacity),
                 _p(bg), _p(r.final_T), _p(r.n_contrib), _p(r.color))
    r._pl_buf = pl_buf
    return r


def backward_blend(r, dL_dpix):
    """Per-tile gradient blend only (backward.cu:428-581): dict with dL_dmean2D [P,2],
    dL_dconic [P,3] (x,y,w), dL_dopacity [P,1], dL_dcolor [P,3]."""
    l = _lib(r.precision)
    real = np.float32 if r.precision == "f32" else np.float64
    P, W, H = r.P, r.W, r.H
    i = r.inputs
    dL_dpix = np.ascontiguousarray(np.asarray(dL_dpix, dtype=real))
    g = dict(dL_dmean2D=np.zeros((P, 2), real), dL_dconic=np.zeros((P, 3), real),
             dL_dopacity=np.zeros((P, 1), real), dL_dcolor=np.zeros((P, 3), real))
    l.gso_render_backward(W, H, _p(r.ranges), _p(r._pl_buf), _p(i["bg"]), _p(r.means2D),
                          _p(r.conic_opacity), _p(r.colors), _p(r.final_T), _p(r.n_contrib),
                          _p(dL_dpix), _p(g["dL_dmean2D"]), _p(g["dL_dconic"]), _p(g["dL_dopacity"]),
                          _p(g["dL_dcolor"]))
    return g


def backward_geometry(r, blend):
    """Per-Gaussian geometry backward (backward.cu:143-293, 378-425) from backward_blend()'s
    (possibly cross-rank reduced) output. Returns the remaining gradient arrays."""
    l = _lib(r.precision)
    real = np.float32 if r.precision == "f32" else np.float64
    P, W, H, M = r.P, r.W, r.H, r.M
    i = r.inputs
    c = lambda a: np.ascontiguousarray(np.asarray(a, dtype=real))
    d2, dc, dcol = c(blend["dL_dmean2D"]), c(blend["dL_dconic"]), c(blend["dL_dcolor"])
    g = dict(dL_dmean3D=np.zeros((P, 3), real), dL_dcov3D=np.zeros((P, 6), real),
             dL_dsh=np.zeros((P, M, 3), real), dL_dscale=np.zeros((P, 3), real),
             dL_drot=np.zeros((P, 4), real))
    has_sh = i["shs"] is not None
    has_scale = i["scales"] is not None and i["cov3D_precomp"] is None
    l.gso_geometry_backward(P, r.D, M, W, H, _p(i["means3D"]), _p(r.radii), _p(i["shs"]),
                            _p(r.clamped), _p(i["scales"]) if has_scale else None,
                            _p(i["rotations"]) if has_scale else None,
                            ctypes.c_float(i["scale_modifier"]), _p(i["cov3D_precomp"]), _p(i["V"]),
                            _p(i["PM"]), _p(i["campos"]), ctypes.c_float(i["tanfovx"]),
                            ctypes.c_float(i["tanfovy"]), _p(d2), _p(dc), _p(dcol),
                            _p(g["dL_dmean3D"]), _p(g["dL_dcov3D"]),
                            _p(g["dL_dsh"]) if has_sh else None,
                            _p(g["dL_dscale"]) if has_scale else None,
                            _p(g["dL_drot"]) if has_scale else None)
    return g


def backward(r, dL_dpix):
    """Gradients for an OracleResult. Returns dict with the reference's 8 output tensors
    (+ dL_dconic [P,3] = (x,y,w))."""
    g = backward_blend(r, dL_dpix)
    g.update(backward_geometry(r, g))
    return g


def forward_scene(s, precision="f32"):
    """Convenience: run on a gaussiancity_b200.synthetic.Scene (tensors may live on any device)."""
    c = lambda t: None if t is None else t.detach().cpu().numpy()
    return forward(c(s.means3D), c(s.opacities), c(s.scales), c(s.rotations), c(s.view_matrix),
                   c(s.proj_matrix), c(s.campos), s.img_w, s.img_h, s.tanfovx, s.tanfovy, c(s.bg),
                   shs=c(s.shs), colors_precomp=c(s.colors_precomp), sh_degree=s.sh_degree,
                   precision=precision)
