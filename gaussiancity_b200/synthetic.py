"""Seeded synthetic Gaussian clouds and cameras for parity tests and bench.py.

Two distributions:
  * "uniform": z ~ U(2,20), x/y spread 1.1x beyond the frustum, log-uniform projected sigma in
    [0.5, 4] px with per-axis jitter, random rotations, opacity ~ U(0.1,1), SH DC ~ N(0,0.5),
    higher bands ~ N(0,0.1).  Camera at the origin looking +z, fovx 60 deg.
  * "city": integer voxel lattice (ground plane + box buildings), identity quaternions,
    opacity 1, scale 0.5, colours_precomp ~ U(-1,1), rendered through the GaussianCity camera
    adapter (negative clip-space w, flips) -- mirrors what utils/helpers.get_gaussian_points
    feeds the rasterizer (reference utils/helpers.py:226-247, config.py:117).
Everything is generated with an explicit torch.Generator so CPU and CUDA callers agree.
"""
import math
from typing import NamedTuple, Optional

import numpy as np
import torch


class Scene(NamedTuple):
    means3D: torch.Tensor          # [P,3]
    scales: torch.Tensor           # [P,3]
    rotations: torch.Tensor        # [P,4] (r,x,y,z)
    opacities: torch.Tensor        # [P,1]
    shs: Optional[torch.Tensor]    # [P,M,3] or None
    colors_precomp: Optional[torch.Tensor]  # [P,3] or None
    sh_degree: int
    img_w: int
    img_h: int
    tanfovx: float
    tanfovy: float
    view_matrix: torch.Tensor      # [4,4] transposed (row-vector convention)
    proj_matrix: torch.Tensor      # [4,4] transposed full world->clip
    campos: torch.Tensor           # [3]
    bg: torch.Tensor               # [3]


def synthetic_camera(img_w, img_h, fovx_deg=60.0, znear=0.01, zfar=100.0, device="cpu"):
    tanfovx = math.tan(math.radians(fovx_deg) * 0.5)
    tanfovy = tanfovx * img_h / img_w
    view = torch.eye(4, dtype=torch.float32)
    Pm = torch.zeros(4, 4, dtype=torch.float32)
    Pm[0, 0] = 1.0 / tanfovx
    Pm[1, 1] = 1.0 / tanfovy
    Pm[2, 2] = zfar / (zfar - znear)
    Pm[2, 3] = -(zfar * znear) / (zfar - znear)
    Pm[3, 2] = 1.0
    proj = (Pm @ view).T.contiguous()
    return (tanfovx, tanfovy, view.T.contiguous().to(device), proj.to(device),
            torch.zeros(3, dtype=torch.float32, device=device))


def uniform_scene(P, img_w, img_h, sh_degree=0, seed=0, device="cpu", use_sh=True,
                  sigma_px=(0.5, 4.0), bg=(0.0, 0.0, 0.0), spread=1.1):
    """SURVEY 8d "uniform" cloud.  `spread` = lateral extent in units of tan(fov/2) (1.1: ~17 % of the
    centres fall outside the image); with spread > 1.3 and large `sigma_px` some splats whose
    view-space x/z, y/z is CLAMPED to +-1.3 tan(fov/2) (forward.cu:80-85, backward.cu x_grad_mul)
    still reach the screen -- the case the default never produces."""
    g = torch.Generator(device="cpu").manual_seed(seed)

    def U(*shape, lo=0.0, hi=1.0):
        return torch.rand(*shape, generator=g, dtype=torch.float32) * (hi - lo) + lo

    def N(*shape, std=1.0):
        return torch.randn(*shape, generator=g, dtype=torch.float32) * std

    tanfovx, tanfovy, view, proj, campos = synthetic_camera(img_w, img_h, device=device)
    focal = img_w / (2.0 * tanfovx)
    z = U(P, lo=2.0, hi=20.0)
    x = U(P, lo=-1.0, hi=1.0) * z * tanfovx * spread
    y = U(P, lo=-1.0, hi=1.0) * z * tanfovy * spread
    means = torch.stack([x, y, z], dim=1)
    lo, hi = math.log(sigma_px[0]), math.log(sigma_px[1])
    sig = torch.exp(U(P, lo=lo, hi=hi))
    base = sig * z / focal
    scales = base[:, None] * U(P, 3, lo=0.5, hi=2.0)
    rot = N(P, 4)
    rot = rot / rot.norm(dim=1, keepdim=True)
    opac = U(P, 1, lo=0.1, hi=1.0)
    M = (sh_degree + 1) ** 2
    shs = colors = None
    if use_sh:
        shs = torch.cat([N(P, 1, 3, std=0.5), N(P, M - 1, 3, std=0.1)], dim=1) if M > 1 else N(P, 1, 3, std=0.5)
        shs = shs.contiguous()
    else:
        colors = U(P, 3, lo=-1.0, hi=1.0)
    to = lambda t: None if t is None else t.to(device)
    return Scene(to(means), to(scales), to(rot), to(opac), to(shs), to(colors), sh_degree, img_w,
                 img_h, tanfovx, tanfovy, view, proj, campos,
                 torch.tensor(bg, dtype=torch.float32, device=device))


# GoogleEarth intrinsics used by the reference (config.py:36-37): 960x540 sensor
CITY_K = np.array([[1528.1469407, 0.0, 480.0], [0.0, 1528.1469407, 270.0], [0.0, 0.0, 1.0]],
                  dtype=np.float64)
CITY_SENSOR = (960, 540)


def city_points(P, seed=0, extent=256, device="cpu"):
    """[P,14] GaussianCity-style points on an integer lattice + an orbit camera pose
    (cam_position[3], cam_quaternion xyzw[4]) looking at the centre from ~600 units."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    n_ground = P // 2
    side = int(math.ceil(math.sqrt(n_ground)))
    gx, gy = torch.meshgrid(torch.arange(side), torch.arange(side), indexing="ij")
    ground = torch.stack([gx.flatten(), gy.flatten(), torch.zeros(side * side)], dim=1)[:n_ground]
    ground = ground.float() * (extent / side)
    ground = torch.floor(ground)
    n_bld = P - n_ground
    bx = torch.randint(0, extent, (n_bld,), generator=g).float()
    by = torch.randint(0, extent, (n_bld,), generator=g).float()
    bz = torch.randint(1, 64, (n_bld,), generator=g).float()
    xyz = torch.cat([ground, torch.stack([bx, by, bz], dim=1)], dim=0)
    xyz[:, 0] -= extent / 2
    xyz[:, 1] -= extent / 2
    opacity = torch.ones(P, 1)
    scales = torch.full((P, 3), 0.5)
    scales[:n_ground, 2] = 1.0
    rot = torch.zeros(P, 4)
    rot[:, 0] = 1.0
    rgb = torch.rand(P, 3, generator=g) * 2 - 1
    pts = torch.cat([xyz, opacity, scales, rot, rgb], dim=1).float().contiguous()
    # orbit pose: camera on a circle of radius ~520 at altitude ~300 looking at the origin
    theta = 0.7
    cam_pos = np.array([520.0 * math.cos(theta), 520.0 * math.sin(theta), 300.0])
    fwd = -cam_pos / np.linalg.norm(cam_pos)
    up0 = np.array([0.0, 0.0, 1.0])
    right = np.cross(up0, fwd)   # F x ... order chosen so that [F|R|U] is right-handed
    right /= np.linalg.norm(right)
    up = np.cross(fwd, right)
    Rm = np.stack([fwd, right, up], axis=1)  # columns F | R | U (reference axis convention)
    quat = _matrix_to_quat_xyzw(Rm)
    return pts.to(device), cam_pos, quat


def _matrix_to_quat_xyzw(R):
    t = np.trace(R)
    if t > 0:
        s = math.sqrt(t + 1.0) * 2
        w = 0.25 * s
        x = (R[2, 1] - R[1, 2]) / s
        y = (R[0, 2] - R[2, 0]) / s
        z = (R[1, 0] - R[0, 1]) / s
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = math.sqrt(max(R[i, i] - R[j, j] - R[k, k] + 1.0, 1e-12)) * 2
        q = [0.0, 0.0, 0.0]
        q[i] = 0.25 * s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
        w = (R[k, j] - R[j, k]) / s
        x, y, z = q
    q = np.array([x, y, z, w])
    return q / np.linalg.norm(q)


def city_scene(P, seed=0, device="cpu", scale=1.0, extent=256):
    """The GaussianCity call pattern as a Scene: lattice points through the K / sensor camera
    adapter (negative clip-space w), colours precomputed in (-1,1), identity quaternions,
    opacity 1 (reference utils/helpers.py:226-247 + DGR/__init__.py:382-426)."""
    from . import GaussianRasterizerWrapper
    K = CITY_K.copy()
    K[:2] *= scale
    sensor = (int(CITY_SENSOR[0] * scale), int(CITY_SENSOR[1] * scale))
    pts, cam_pos, cam_quat = city_points(P, seed=seed, extent=extent, device=device)
    wrap = GaussianRasterizerWrapper(K, sensor, device=torch.device(device) if isinstance(device, str) else device)
    st = wrap._get_gaussian_rasterization_settings(cam_pos, cam_quat)
    return Scene(pts[:, 0:3].contiguous(), pts[:, 4:7].contiguous(), pts[:, 7:11].contiguous(),
                 pts[:, 3:4].contiguous(), None, pts[:, 11:14].contiguous(), 0, st.img_w, st.img_h,
                 st.tanfovx, st.tanfovy, st.view_matrix.contiguous(), st.proj_matrix.contiguous(),
                 st.campos.contiguous(), st.bg)
