"""Generate golden vectors from the UNMODIFIED reference extension (oracle/_ref) -- run on the
B200 box via gpurun:  python tools/make_golden.py gpurun_out/golden
The .npz files are then committed under tests/golden/ and pin the CPU oracle
(tests/test_oracle_golden.py) without needing /root/reference or a GPU."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gaussiancity_b200 import GaussianRasterizerWrapper  # noqa: E402  (camera math only)
from gaussiancity_b200.synthetic import (CITY_K, CITY_SENSOR, Scene, city_points,  # noqa: E402
                                         uniform_scene)
from tests import refext  # noqa: E402


def city_scene(P, seed, device, scale=0.25):
    """GaussianCity call pattern through the camera adapter: negative clip-space w, colours
    precomputed in (-1,1), identity quaternions, opacity 1. Sensor scaled down for a small file."""
    K = CITY_K.copy()
    K[:2] *= scale
    sensor = (int(CITY_SENSOR[0] * scale), int(CITY_SENSOR[1] * scale))
    pts, cam_pos, cam_quat = city_points(P, seed=seed, extent=96, device=device)
    wrap = GaussianRasterizerWrapper(K, sensor, device=device)
    st = wrap._get_gaussian_rasterization_settings(cam_pos, cam_quat)
    return Scene(pts[:, 0:3].contiguous(), pts[:, 4:7].contiguous(), pts[:, 7:11].contiguous(),
                 pts[:, 3:4].contiguous(), None, pts[:, 11:14].contiguous(), 0, st.img_w, st.img_h,
                 st.tanfovx, st.tanfovy, st.view_matrix.contiguous(), st.proj_matrix.contiguous(),
                 st.campos.contiguous(), st.bg)


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    ref = refext.load_reference_ext()
    assert ref is not None, "oracle/_ref is not built"
    dev = torch.device("cuda:0")
    cases = {
        "uniform_sh3_400_96x64": uniform_scene(400, 96, 64, sh_degree=3, seed=21, device=dev, bg=(0.1, 0.3, 0.2)),
        "uniform_sh1_300_70x50": uniform_scene(300, 70, 50, sh_degree=1, seed=22, device=dev),
        "uniform_sh0_500_80x80": uniform_scene(500, 80, 80, sh_degree=0, seed=23, device=dev, bg=(1.0, 1.0, 1.0)),
        "uniform_precomp_300_64x48": uniform_scene(300, 64, 48, seed=24, device=dev, use_sh=False),
        "city_wrapper_600": city_scene(600, 25, dev),
    }
    for name, s in cases.items():
        args = refext.scene_forward_args(s)
        R, color, radii, geom, binning, img = ref.rasterize_gaussians(*args)
        g = torch.Generator().manual_seed(99)
        grad_out = torch.randn(3, s.img_h, s.img_w, generator=g).to(dev)
        grads = ref.rasterize_gaussians_backward(*refext.scene_backward_args(s, radii, grad_out, geom, R, binning, img))
        torch.cuda.synchronize()
        P, W, H = s.means3D.shape[0], s.img_w, s.img_h
        gv = refext.ref_geom_views(geom, P)
        iv = refext.ref_img_views(img, H, W)
        tiles = ((W + 15) // 16) * ((H + 15) // 16)
        out = dict(
            means3D=s.means3D, scales=s.scales, rotations=s.rotations, opacities=s.opacities,
            view_matrix=s.view_matrix, proj_matrix=s.proj_matrix, campos=s.campos, bg=s.bg,
            img_w=np.int32(W), img_h=np.int32(H), tanfovx=np.float64(s.tanfovx),
            tanfovy=np.float64(s.tanfovy), sh_degree=np.int32(s.sh_degree), grad_out=grad_out,
            num_rendered=np.int64(R), color=color, radii=radii,
            depths=gv["depths"], means2D=gv["means2D"], cov3D=gv["cov3D"],
            conic_opacity=gv["conic_opacity"], rgb=gv["rgb"], clamped=gv["clamped"].to(torch.uint8),
            tiles_touched=gv["tiles_touched"], final_T=iv["accum_alpha"], n_contrib=iv["n_contrib"],
            ranges=iv["ranges"][:tiles],
            dL_dmeans2D=grads[0], dL_dcolors=grads[1], dL_dopacity=grads[2], dL_dmeans3D=grads[3],
            dL_dcov3D=grads[4], dL_dsh=grads[5], dL_dscales=grads[6], dL_drotations=grads[7])
        if s.shs is not None:
            out["shs"] = s.shs
        if s.colors_precomp is not None:
            out["colors_precomp"] = s.colors_precomp
        if R > 0:
            bv = refext.ref_binning_views(binning, R)
            out["point_list"] = bv["point_list"]
            out["point_list_keys"] = bv["point_list_keys"]
        npd = {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in out.items()}
        path = os.path.join(outdir, name + ".npz")
        np.savez_compressed(path, **npd)
        print(name, "P", P, "R", R, "visible", int((radii > 0).sum()), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
