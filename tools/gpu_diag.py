"""Diagnostic (GPU box): per-quantity mismatch statistics between ours and the reference ext."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussiancity_b200 import ext as ours, _cabi
from gaussiancity_b200.synthetic import uniform_scene
from tests import refext

ref = refext.load_reference_ext()
dev = torch.device("cuda:0")

def ulp(a, b):
    return (a.view(torch.int32).long() - b.view(torch.int32).long()).abs()

def report(name, a, b):
    ne = (a != b)
    n = ne.sum().item()
    msg = f"  {name:28s} mismatches {n:9d} / {a.numel()}"
    if n and a.dtype == torch.float32:
        u = ulp(a[ne], b[ne])
        msg += f"  max_ulp {u.max().item()}  max_abs {(a[ne]-b[ne]).abs().max().item():.3e}"
    print(msg)

for (P, W, H, deg, use_sh, seed) in [(1000,128,128,0,False,0), (100000,512,512,3,True,1), (300000, 1920, 1080, 0, True, 2)]:
    print(f"=== P={P} {W}x{H} deg={deg} sh={use_sh}")
    s = uniform_scene(P, W, H, sh_degree=deg, seed=seed, device=dev, use_sh=use_sh, bg=(0.1,0.2,0.3))
    args = refext.scene_forward_args(s)
    R_ref, col_ref, radii_ref, geom_ref, bin_ref, img_ref = ref.rasterize_gaussians(*args)
    cov = torch.zeros(P, 6, device=dev)
    _cabi.lib().gcr_debug_set_cov3d_out(ctypes.c_void_p(cov.data_ptr()))
    R, col, radii, geom, binning, img = ours.rasterize_gaussians(*args)
    torch.cuda.synchronize()
    print("  R", R, R_ref)
    gv = refext.ref_geom_views(geom_ref, P); ov = refext.our_views(P, R, W, H, geom, binning, img)
    vis = radii_ref > 0
    report("radii", radii, radii_ref)
    report("tiles_touched", ov["tiles_touched"], gv["tiles_touched"])
    report("cov3D", cov[vis], gv["cov3D"][vis])
    rec = ov["records"]
    report("mean2D", rec[vis][:, 0:2], gv["means2D"][vis])
    for k, nm in enumerate(["conic.x", "conic.y"]):
        report(nm, rec[vis][:, 2 + k], gv["conic_opacity"][vis][:, k])
    report("conic.z", rec[vis][:, 4], gv["conic_opacity"][vis][:, 2])
    report("opacity", rec[vis][:, 5], gv["conic_opacity"][vis][:, 3])
    if use_sh:
        report("rgb", rec[vis][:, 8:11], gv["rgb"][vis])
    if R == R_ref and R > 0:
        bv = refext.ref_binning_views(bin_ref, R)
        report("point_list", ov["point_list"], bv["point_list"])
    iv = refext.ref_img_views(img_ref, H, W)
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    if R == R_ref:
        report("ranges", ov["ranges"], iv["ranges"][:tiles])
    report("n_contrib", ov["n_contrib"], iv["n_contrib"])
    report("final_T", ov["final_T"], iv["accum_alpha"])
    report("color", col, col_ref)
