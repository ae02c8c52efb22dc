"""Adapter fusion: GaussianCity's `utils.helpers.get_gaussian_points` +
`get_gaussian_rasterization` (utils/helpers.py:226-270) without their intermediate tensors.

The reference builds a `[B, N, 14]` tensor per step -- `cat(xyz, opacity, scales, rotations, rgb)`
with `ones` for the opacity and identity quaternions for the rotations whenever the generator does
not predict them (its shipped configurations predict `rgb` only, config.py:128) -- slices it apart
again in the wrapper (DGR/__init__.py:404-417), renders the full 960x540 frame per batch element
and then crops 640x448 out of it.  Here the generator's attribute tensors go to the kernels as
they are (structure of arrays), a missing opacity / rotation is a NULL pointer the projection
kernel reads as 1 / identity (bit-identical arithmetic, csrc/preprocess.cu), and the crop is a
pixel window inside the rasterizer: tiles outside it are neither binned nor blended, the output is
the cropped image, and every pixel of it has exactly the value the full render would give it
(the tile grid and the camera stay those of the full frame -- DESIGN.md section 8 explains why a
shifted camera would not be result-preserving).
"""
import torch

from . import ext

__all__ = ["render_gaussian_points", "get_gaussian_rasterization_fused", "crop_to_window"]


def _native():
    """The native torch module when it is built (same C ABI underneath, ~10x less host time per call
    than ctypes -- this path exists for GaussianCity's launch-bound <= 16 k-point frames)."""
    try:
        from .compat import diff_gaussian_rasterization_ext as m
        return m if hasattr(m, "rasterize_gaussians_window") else None
    except ImportError:
        return None


class _WindowRasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, scales, rgb, opacity, rotations, settings, window):
        rs = settings
        nat = _native()
        e = torch.Tensor([])
        if nat is not None:
            R, color, radii, geom, binning, img = nat.rasterize_gaussians_window(
                rs.bg, xyz, rgb, opacity if opacity is not None else e, scales,
                rotations if rotations is not None else e, rs.scale_modifier, rs.view_matrix, rs.proj_matrix,
                rs.tanfovx, rs.tanfovy, rs.img_h, rs.img_w, *window, rs.debug)
        else:
            R, color, radii, geom, binning, img = ext.rasterize_gaussians_window(
                rs.bg, xyz, rgb, opacity, scales, rotations, rs.scale_modifier, rs.view_matrix, rs.proj_matrix,
                rs.tanfovx, rs.tanfovy, rs.img_h, rs.img_w, window, rs.debug)
        ctx.rs, ctx.window, ctx.R = rs, window, R
        ctx.has_opacity, ctx.has_rot = opacity is not None, rotations is not None
        ctx.save_for_backward(xyz, scales, rotations if rotations is not None else e, radii, geom, binning, img)
        return color

    @staticmethod
    def backward(ctx, grad_color):
        rs = ctx.rs
        xyz, scales, rotations, radii, geom, binning, img = ctx.saved_tensors
        nat = _native()
        if nat is not None:
            d_xyz, d_rgb, d_scales, d_opacity, d_rot, _ = nat.rasterize_gaussians_backward_window(
                rs.bg, xyz, radii, scales, rotations, rs.scale_modifier, rs.view_matrix, rs.proj_matrix, rs.tanfovx,
                rs.tanfovy, rs.img_h, rs.img_w, *ctx.window, grad_color, geom, ctx.R, binning, img, ctx.has_opacity,
                rs.debug)
            d_opacity = d_opacity if ctx.has_opacity else None
            d_rot = d_rot if ctx.has_rot else None
        else:
            d_xyz, d_rgb, d_scales, d_opacity, d_rot, _ = ext.rasterize_gaussians_backward_window(
                rs.bg, xyz, radii, scales, rotations if ctx.has_rot else None, rs.scale_modifier, rs.view_matrix,
                rs.proj_matrix, rs.tanfovx, rs.tanfovy, rs.img_h, rs.img_w, ctx.window, grad_color, geom, ctx.R,
                binning, img, want_opacity=ctx.has_opacity, debug=rs.debug)
        return d_xyz, d_scales, d_rgb, d_opacity, d_rot, None, None


def crop_to_window(W, H, crop, flip_lr, flip_ud):
    """Pixel window (x, y, w, h) of the UNFLIPPED render that ends up as `crop` (dict x/y/w/h in the
    coordinates of the wrapper's output, i.e. after its flips; None = everything) once the window is
    flipped the way the wrapper flips the frame (DGR/__init__.py:421-424)."""
    if crop is None:
        x, y, w, h = 0, 0, int(W), int(H)
    else:
        x, y, w, h = int(crop["x"]), int(crop["y"]), int(crop["w"]), int(crop["h"])
    if w <= 0 or h <= 0 or x < 0 or y < 0 or x + w > W or y + h > H:
        raise ValueError(f"crop {crop} does not lie inside the {W}x{H} frame")
    if flip_lr:
        x = int(W) - (x + w)
    if flip_ud:
        y = int(H) - (y + h)
    return x, y, w, h


def render_gaussian_points(xyz, scales, rgb, wrapper, cam_position, cam_quaternion, opacity=None,
                           rotations=None, crop=None):
    """One frame: xyz/scales/rgb [N,3] (+ optional opacity [N,1], rotations [N,4]) through
    `wrapper` (a gaussiancity_b200.GaussianRasterizerWrapper: camera, flips) -> image [3,h,w].
    crop = dict(x=, y=, w=, h=) in the coordinates of the wrapper's OUTPUT image (after its flips),
    exactly like the slice in utils/helpers.py:261-267; None = the full frame."""
    rs = wrapper._get_gaussian_rasterization_settings(cam_position, cam_quaternion)
    x, y, w, h = crop_to_window(rs.img_w, rs.img_h, crop, wrapper.flip_lr, wrapper.flip_ud)
    img = _WindowRasterize.apply(xyz, scales, rgb, opacity, rotations, rs, (x, y, w, h))
    if wrapper.flip_lr:
        img = torch.flip(img, dims=[2])
    if wrapper.flip_ud:
        img = torch.flip(img, dims=[1])
    return img


def get_gaussian_rasterization_fused(xyz, scales, attrs, wrapper, cam_pos, cam_quat, crop_bboxes=None):
    """Drop-in for get_gaussian_points + get_gaussian_rasterization (utils/helpers.py:226-270):
    xyz [B,N,3], scales [B,N,3], attrs = the generator's output dict (rgb; optionally xyz offsets,
    scale factors, opacity) -> images [B,3,h,w].  Unlike the reference this does not modify xyz /
    scales in place."""
    rgb = attrs["rgb"]
    if "xyz" in attrs:
        xyz = xyz + attrs["xyz"]
    if "scale" in attrs:
        scales = scales * attrs["scale"]
    opacity = attrs.get("opacity")
    images = []
    for i in range(xyz.size(0)):
        images.append(render_gaussian_points(
            xyz[i], scales[i], rgb[i], wrapper, cam_pos[i], cam_quat[i],
            opacity=None if opacity is None else opacity[i],
            crop=None if crop_bboxes is None else crop_bboxes[i]))
    return torch.stack(images, dim=0)
