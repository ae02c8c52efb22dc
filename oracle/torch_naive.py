"""Naive pure-PyTorch CPU restatement of the forward (BASELINE config 1: "1k synthetic Gaussians ->
128x128, forward only, naive PyTorch CPU per-pixel blend").  TEST INFRASTRUCTURE / reported CPU
baseline only -- never imported by gaussiancity_b200.

Restates DGR/cuda_rasterizer/forward.cu:147-346 without tiles: per-Gaussian preprocessing is
vectorised over Gaussians, then Gaussians are visited in (depth, index) order and every pixel is
blended at once (a Python loop over Gaussians, tensor ops over the H x W pixels).  Without the tile
bounding squares of the reference a splat also reaches pixels outside its 3-sigma square, so the
per-pixel restriction to the reference's tile rectangle is applied explicitly (rect mask).
fp32 throughout; SH degree 0-3 or precomputed colours.
"""
import math

import torch

C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
      -0.4570457994644658, 1.445305721320277, -0.5900435899266435)


def _sh_to_rgb(deg, shs, dirs):
    x, y, z = dirs[:, 0:1], dirs[:, 1:2], dirs[:, 2:3]
    res = C0 * shs[:, 0]
    if deg > 0:
        res = res - C1 * y * shs[:, 1] + C1 * z * shs[:, 2] - C1 * x * shs[:, 3]
        if deg > 1:
            xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
            res = (res + C2[0] * xy * shs[:, 4] + C2[1] * yz * shs[:, 5] + C2[2] * (2 * zz - xx - yy) * shs[:, 6]
                   + C2[3] * xz * shs[:, 7] + C2[4] * (xx - yy) * shs[:, 8])
            if deg > 2:
                res = (res + C3[0] * y * (3 * xx - yy) * shs[:, 9] + C3[1] * xy * z * shs[:, 10]
                       + C3[2] * y * (4 * zz - xx - yy) * shs[:, 11] + C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * shs[:, 12]
                       + C3[4] * x * (4 * zz - xx - yy) * shs[:, 13] + C3[5] * z * (xx - yy) * shs[:, 14]
                       + C3[6] * x * (xx - 3 * yy) * shs[:, 15])
    return torch.clamp_min(res + 0.5, 0.0)


def _preprocess(s, dt, cov3D=None, reference_clamp_gradient=True):
    """Per-Gaussian forward quantities (DGR forward.cu:147-233), vectorised over Gaussians, in dtype
    `dt`; differentiable w.r.t. any input tensor of `s` that requires grad."""
    P, W, H = s.means3D.shape[0], s.img_w, s.img_h
    V, PM = s.view_matrix.to(dt), s.proj_matrix.to(dt)        # transposed (row-vector) matrices
    p = s.means3D.to(dt)
    ph = torch.cat([p, torch.ones(P, 1, dtype=dt)], dim=1)
    view = ph @ V                                             # [P,4] camera space
    hom = ph @ PM
    pw = 1.0 / (hom[:, 3] + 1e-7)
    ndc = hom[:, :2] * pw[:, None]
    tz = view[:, 2]
    visible = tz > 0.2
    # 3D covariance  Sigma = R S^2 R^T with the UN-normalised quaternion (r, x, y, z)
    q = s.rotations.to(dt)
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.stack([
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)], dim=1),
        torch.stack([2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)], dim=1),
        torch.stack([2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=1)], dim=1)
    Sg = R @ torch.diag_embed(s.scales.to(dt) ** 2) @ R.transpose(1, 2)
    if cov3D is not None:     # precomputed covariance path: [P,6] = (S00,S01,S02,S11,S12,S22), forward.cu:138-143
        c6 = cov3D.to(dt)
        Sg = torch.stack([torch.stack([c6[:, 0], c6[:, 1], c6[:, 2]], dim=1),
                          torch.stack([c6[:, 1], c6[:, 3], c6[:, 4]], dim=1),
                          torch.stack([c6[:, 2], c6[:, 4], c6[:, 5]], dim=1)], dim=1)
    # EWA projection
    fx, fy = W / (2.0 * s.tanfovx), H / (2.0 * s.tanfovy)
    limx, limy = 1.3 * s.tanfovx, 1.3 * s.tanfovy
    tzs = torch.where(visible, tz, torch.ones_like(tz))
    tx = torch.clamp(view[:, 0] / tzs, -limx, limx) * tzs
    ty = torch.clamp(view[:, 1] / tzs, -limy, limy) * tzs
    if reference_clamp_gradient:
        # backward.cu:170-171,278-279 (x_grad_mul / y_grad_mul): a clamped t.x, t.y is a CONSTANT for
        # the reference's backward -- not even its dependence on t.z (= +-lim * t.z) is propagated
        inx = (view[:, 0] / tzs).detach().abs() <= limx
        iny = (view[:, 1] / tzs).detach().abs() <= limy
        tx = torch.where(inx, view[:, 0], tx.detach())
        ty = torch.where(iny, view[:, 1], ty.detach())
    zero = torch.zeros_like(tzs)
    J = torch.stack([torch.stack([fx / tzs, zero, -fx * tx / (tzs * tzs)], dim=1),
                     torch.stack([zero, fy / tzs, -fy * ty / (tzs * tzs)], dim=1)], dim=1)   # [P,2,3]
    Rw = V[:3, :3].T                                          # world -> camera rotation
    T = J @ Rw
    cov = T @ Sg @ T.transpose(1, 2)
    a, b, c = cov[:, 0, 0] + 0.3, cov[:, 0, 1], cov[:, 1, 1] + 0.3
    det = a * c - b * b
    ok = visible & (det != 0)
    dets = torch.where(ok, det, torch.ones_like(det))
    conic = torch.stack([c / dets, -b / dets, a / dets], dim=1)
    mid = 0.5 * (a + c)
    lam = mid + torch.sqrt(torch.clamp_min(mid * mid - det, 0.1))
    radius = torch.ceil(3.0 * torch.sqrt(torch.where(ok, lam, torch.ones_like(lam)))).detach()
    px = ((ndc[:, 0].double() + 1.0) * W - 1.0) * 0.5
    py = ((ndc[:, 1].double() + 1.0) * H - 1.0) * 0.5
    px, py = px.to(dt), py.to(dt)
    gx, gy = (W + 15) // 16, (H + 15) // 16
    pxd, pyd = px.detach(), py.detach()
    x0 = torch.clamp(torch.trunc((pxd - radius) / 16), 0, gx)
    y0 = torch.clamp(torch.trunc((pyd - radius) / 16), 0, gy)
    x1 = torch.clamp(torch.trunc((pxd + radius + 15) / 16), 0, gx)
    y1 = torch.clamp(torch.trunc((pyd + radius + 15) / 16), 0, gy)
    ok = ok & ((x1 - x0) * (y1 - y0) > 0)
    radii = torch.where(ok, radius, torch.zeros_like(radius)).to(torch.int32)
    if s.colors_precomp is not None:
        rgb = s.colors_precomp.to(dt)
    else:
        d = p - s.campos.to(dt)[None]
        rgb = _sh_to_rgb(s.sh_degree, s.shs.to(dt), d / d.norm(dim=1, keepdim=True))
    opac = s.opacities.to(dt).reshape(-1)
    idx = torch.nonzero(ok).flatten()
    order = idx[torch.sort(tz.detach()[idx], stable=True).indices]
    return dict(ndc=ndc, px=px, py=py, conic=conic, rgb=rgb, opac=opac, radii=radii, order=order,
                rect=(x0, x1, y0, y1))


def render_naive(s):
    """s: gaussiancity_b200.synthetic.Scene with CPU tensors -> (color[3,H,W], radii[P]); fp32."""
    f32 = torch.float32
    W, H = s.img_w, s.img_h
    with torch.no_grad():
        q = _preprocess(s, f32)
    px, py, conic, rgb, opac = q["px"], q["py"], q["conic"], q["rgb"], q["opac"]
    x0, x1, y0, y1 = q["rect"]
    # front-to-back over Gaussians in (depth, index) order, all pixels of the splat's rectangle at once
    ys, xs = torch.meshgrid(torch.arange(H, dtype=f32), torch.arange(W, dtype=f32), indexing="ij")
    Tr = torch.ones(H, W)
    done = torch.zeros(H, W, dtype=torch.bool)
    C = torch.zeros(3, H, W)
    for g in q["order"].tolist():
        rx0, rx1, ry0, ry1 = int(x0[g]) * 16, int(x1[g]) * 16, int(y0[g]) * 16, int(y1[g]) * 16
        sl = (slice(ry0, min(ry1, H)), slice(rx0, min(rx1, W)))       # the reference's tile rectangle
        dx, dy = px[g] - xs[sl], py[g] - ys[sl]
        power = -0.5 * (conic[g, 0] * dx * dx + conic[g, 2] * dy * dy) - conic[g, 1] * dx * dy
        alpha = torch.clamp_max(opac[g] * torch.exp(power), 0.99)
        contrib = (~done[sl]) & (power <= 0) & (alpha >= 1.0 / 255.0)
        test_T = Tr[sl] * (1 - alpha)
        sat = contrib & (test_T < 1e-4)
        done[sl] |= sat
        upd = contrib & ~sat
        w = torch.where(upd, alpha * Tr[sl], torch.zeros_like(alpha))
        C[:, sl[0], sl[1]] += rgb[g][:, None, None] * w[None]
        Tr[sl] = torch.where(upd, test_T, Tr[sl])
    return C + Tr[None] * s.bg.to(f32)[:, None, None], q["radii"]


def render_autograd(s, dtype=torch.float64, reference_cap_gradient=True, cov3D=None,
                    reference_clamp_gradient=True):
    """Differentiable restatement for gradient ground truth (SURVEY 8c "fp64 + autograd"): the same
    forward written functionally (no in-place updates), so torch.autograd differentiates it.  The
    thresholds (alpha < 1/255, T < 1e-4, the frustum clamp of t, the SH clamp at 0) are piecewise
    constant exactly as the reference's analytic backward treats them (DGR backward.cu).  One
    reference quirk is reproduced on request (default): backward.cu:523,562,578 recomputes
    `alpha = min(0.99, o G)` but propagates `dL/dalpha` to G and o as if the cap were not there, so
    with `reference_cap_gradient=True` the cap is applied straight-through (capped value, identity
    gradient); with False the mathematically exact derivative (zero where capped) is taken -- the two
    differ wherever o G > 0.99, i.e. all over GaussianCity's opacity-1 scenes.  Likewise
    `reference_clamp_gradient` (see _preprocess): a frustum-clamped t.x / t.y is a constant for the
    reference's backward.  Returns (color[3,H,W], ndc[P,2]); `ndc` is the retained intermediate whose
    gradient is the reference's `dL_dmeans2D` ("viewspace points").  Small scenes only: every
    Gaussian touches a full-image tensor."""
    W, H = s.img_w, s.img_h
    q = _preprocess(s, dtype, cov3D=cov3D, reference_clamp_gradient=reference_clamp_gradient)
    if q["ndc"].requires_grad:
        q["ndc"].retain_grad()
    px, py, conic, rgb, opac = q["px"], q["py"], q["conic"], q["rgb"], q["opac"]
    x0, x1, y0, y1 = q["rect"]
    ys, xs = torch.meshgrid(torch.arange(H, dtype=dtype), torch.arange(W, dtype=dtype), indexing="ij")
    Tr = torch.ones(H, W, dtype=dtype)
    done = torch.zeros(H, W, dtype=torch.bool)
    C = torch.zeros(3, H, W, dtype=dtype)
    for g in q["order"].tolist():
        inrect = ((xs >= int(x0[g]) * 16) & (xs < int(x1[g]) * 16) &
                  (ys >= int(y0[g]) * 16) & (ys < int(y1[g]) * 16))
        dx, dy = px[g] - xs, py[g] - ys
        power = -0.5 * (conic[g, 0] * dx * dx + conic[g, 2] * dy * dy) - conic[g, 1] * dx * dy
        raw = opac[g] * torch.exp(torch.clamp_max(power, 0.0))
        alpha = torch.clamp_max(raw, 0.99)
        if reference_cap_gradient:
            alpha = raw + (alpha - raw).detach()
        contrib = inrect & (~done) & (power.detach() <= 0) & (alpha.detach() >= 1.0 / 255.0)
        test_T = Tr * (1 - alpha)
        sat = contrib & (test_T.detach() < 1e-4)
        done = done | sat
        upd = contrib & ~sat
        C = C + rgb[g][:, None, None] * torch.where(upd, alpha * Tr, torch.zeros_like(alpha))[None]
        Tr = torch.where(upd, test_T, Tr)
    return C + Tr[None] * s.bg.to(dtype)[:, None, None], q["ndc"]
