"""Tiny forward+backward workloads for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussiancity_b200 import ext
from gaussiancity_b200.synthetic import uniform_scene
from tests import refext

dev = torch.device("cuda:0")
for (P, W, H, deg, use_sh, sig) in [(3000, 96, 80, 3, True, (0.5, 4.0)), (1500, 50, 33, 0, False, (3.0, 12.0)), (40, 16, 16, 1, True, (0.5, 2.0))]:
    s = uniform_scene(P, W, H, sh_degree=deg, seed=5, device=dev, use_sh=use_sh, sigma_px=sig)
    R, color, radii, geom, binning, img = ext.rasterize_gaussians(*refext.scene_forward_args(s))
    G = torch.ones(3, H, W, device=dev)
    grads = ext.rasterize_gaussians_backward(*refext.scene_backward_args(s, radii, G, geom, R, binning, img))
    for k in range(2):   # striped variant too: forward + both backward halves
        Rk, ck, rk, gk, bk, ik = ext.rasterize_gaussians(*refext.scene_forward_args(s), shard_rank=k, shard_count=2)
        acc = ext.rasterize_gaussians_backward_blend(s.bg, P, Rk, G, gk, bk, ik, shard_rank=k, shard_count=2)
        e = torch.Tensor([])
        ext.rasterize_gaussians_backward_geometry(s.means3D, rk, s.scales, s.rotations, 1.0, e, s.view_matrix,
            s.proj_matrix, s.tanfovx, s.tanfovy, H, W, s.shs if s.shs is not None else e, s.sh_degree, s.campos,
            gk, acc, shard_rank=k, striped=True)
    if not use_sh:   # crop folded into the rasterizer, NULL opacity / rotation (adapter path)
        win = (3, 5, W - 9, H - 7)
        Rw, cw, rw, gw, bw, iw = ext.rasterize_gaussians_window(s.bg, s.means3D, s.colors_precomp, None, s.scales, None, 1.0,
                                                                s.view_matrix, s.proj_matrix, s.tanfovx, s.tanfovy, H, W, win)
        ext.rasterize_gaussians_backward_window(s.bg, s.means3D, rw, s.scales, None, 1.0, s.view_matrix, s.proj_matrix,
                                                s.tanfovx, s.tanfovy, H, W, win, torch.ones_like(cw), gw, Rw, bw, iw)
    torch.cuda.synchronize()
    print("ok", P, W, H, "R", R, float(color.sum()), float(grads[3].abs().sum()))
