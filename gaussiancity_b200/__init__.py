"""gaussiancity_b200 -- B200-native differentiable Gaussian rasterizer, drop-in for
hzxie/GaussianCity's extensions/diff_gaussian_rasterization ("DGR").

Seam B (this module) mirrors DGR/__init__.py name for name:
  RasterizeGaussiansFunction   DGR/__init__.py:19-200
  GaussianRasterizationSettings DGR/__init__.py:203-215
  GaussianRasterizer           DGR/__init__.py:218-273
  GaussianRasterizerWrapper    DGR/__init__.py:276-426
so `utils/helpers.get_gaussian_rasterization` / `core/train.py` can switch with one import.
Seam A (`gaussiancity_b200.ext`, also importable as `diff_gaussian_rasterization_ext` from
`gaussiancity_b200/compat`) provides the three native functions, over the C ABI in
include/gcr_rasterizer.h.  The hot path is hand-written CUDA for sm_100a; there is no CPU path.
"""
import math
import typing

import numpy as np
import torch

from . import ext as _ctypes_ext
from .camera import PoseSettingsCache, w2c_host

# Host binding of the three native entry points.  Both bindings are thin layers over the same C
# ABI (include/gcr_rasterizer.h) and the same CUDA library -- there is no CPU path behind either:
#   compat/diff_gaussian_rasterization_ext  native torch module (csrc/torch_module.cpp): what the
#       reference's own Python imports (Seam A); ~10 us of host time per call, which is what counts
#       in GaussianCity's own regime (<= 16 384 points per frame, launch/host bound);
#   ext.py  ctypes: no compiler needed beyond nvcc, carries the striped (multi-GPU) extensions.
# GCR_HOST_BINDING=ctypes|native forces one (native raises if it has not been built).
import os as _os


def _pick_binding():
    want = _os.environ.get("GCR_HOST_BINDING", "")
    if want == "ctypes":
        return _ctypes_ext, "ctypes"
    try:
        from .compat import diff_gaussian_rasterization_ext as native
        if native.abi_version() == _ctypes_ext._cabi.ABI_VERSION:
            return native, "native"
        if want == "native":
            raise ImportError("native module was built against another ABI version: rebuild")
    except ImportError:
        if want == "native":
            raise
    return _ctypes_ext, "ctypes"


dgr_ext, HOST_BINDING = _pick_binding()

__all__ = [
    "RasterizeGaussiansFunction",
    "GaussianRasterizationSettings",
    "GaussianRasterizer",
    "GaussianRasterizerWrapper",
    "mark_visible",
]


class GaussianRasterizationSettings(typing.NamedTuple):
    img_h: int
    img_w: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    view_matrix: torch.Tensor
    proj_matrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


def _cpu_snapshot(args):
    return tuple(a.detach().cpu().clone() if isinstance(a, torch.Tensor) else a for a in args)


class RasterizeGaussiansFunction(torch.autograd.Function):
    """autograd bridge; argument order and saved state follow DGR/__init__.py:29-200."""

    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                cov3Ds_precomp, raster_settings):
        rs = raster_settings
        args = (rs.bg, means3D, colors_precomp, opacities, scales, rotations, rs.scale_modifier,
                cov3Ds_precomp, rs.view_matrix, rs.proj_matrix, rs.tanfovx, rs.tanfovy, rs.img_h,
                rs.img_w, sh, rs.sh_degree, rs.campos, rs.prefiltered, rs.debug)
        if rs.debug:
            snapshot = _cpu_snapshot(args)  # before anything can corrupt them
            try:
                out = dgr_ext.rasterize_gaussians(*args)
            except Exception:
                torch.save(snapshot, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise
        else:
            out = dgr_ext.rasterize_gaussians(*args)
        num_rendered, color, radii, geom_buffer, binning_buffer, img_buffer = out

        ctx.raster_settings = rs
        ctx.num_rendered = num_rendered
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh,
                              geom_buffer, binning_buffer, img_buffer)
        ctx.mark_non_differentiable(radii)
        return color, radii

    @staticmethod
    def backward(ctx, grad_out_color, _grad_radii):
        rs = ctx.raster_settings
        (colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geom_buffer,
         binning_buffer, img_buffer) = ctx.saved_tensors
        args = (rs.bg, means3D, radii, colors_precomp, scales, rotations, rs.scale_modifier,
                cov3Ds_precomp, rs.view_matrix, rs.proj_matrix, rs.tanfovx, rs.tanfovy,
                grad_out_color, sh, rs.sh_degree, rs.campos, geom_buffer, ctx.num_rendered,
                binning_buffer, img_buffer, rs.debug)
        if rs.debug:
            snapshot = _cpu_snapshot(args)
            try:
                grads = dgr_ext.rasterize_gaussians_backward(*args)
            except Exception:
                torch.save(snapshot, "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                raise
        else:
            grads = dgr_ext.rasterize_gaussians_backward(*args)
        (grad_means2D, grad_colors_precomp, grad_opacities, grad_means3D, grad_cov3Ds_precomp,
         grad_sh, grad_scales, grad_rotations) = grads
        # input order of forward(); None for raster_settings
        return (grad_means3D, grad_means2D, grad_sh, grad_colors_precomp, grad_opacities,
                grad_scales, grad_rotations, grad_cov3Ds_precomp, None)


class GaussianRasterizer(torch.nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        with torch.no_grad():
            rs = self.raster_settings
            return dgr_ext.mark_visible(positions, rs.view_matrix, rs.proj_matrix)

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None,
                rotations=None, cov3D_precomp=None):
        if (shs is None) == (colors_precomp is None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        has_sr = scales is not None or rotations is not None
        if ((scales is None or rotations is None) and cov3D_precomp is None) or (
                has_sr and cov3D_precomp is not None):
            raise Exception(
                "Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        empty = torch.Tensor([])
        shs = empty if shs is None else shs
        colors_precomp = empty if colors_precomp is None else colors_precomp
        scales = empty if scales is None else scales
        rotations = empty if rotations is None else rotations
        cov3D_precomp = empty if cov3D_precomp is None else cov3D_precomp
        return RasterizeGaussiansFunction.apply(means3D, means2D, shs, colors_precomp, opacities,
                                                scales, rotations, cov3D_precomp,
                                                self.raster_settings)


class GaussianRasterizerWrapper(torch.nn.Module):
    """GaussianCity's camera adapter (DGR/__init__.py:276-426): intrinsics K + sensor size give
    the FoV and an OpenGL-style projection with P[3,2] = -1; a pose (position, xyzw quaternion)
    gives w2c with the axis permutation [F|R|U] -> [R|U|F]; points are [N,14] =
    xyz | opacity | scale | rotation | rgb; the image is flipped along W by default.

    `fast_camera=True` (not in the reference) builds the per-pose settings on the host and uploads
    them in one packed copy with an LRU pose cache (gaussiancity_b200/camera.py)
    instead of the reference's device matmul + device 4x4 inverse + three small copies per call."""

    def __init__(self, K, sensor_size, flip_lr=True, flip_ud=False, z_near=0.01, z_far=50000.0,
                 device=torch.device("cuda"), fast_camera=False):
        super().__init__()
        self.flip_lr, self.flip_ud = flip_lr, flip_ud
        self.z_near, self.z_far = z_near, z_far
        self.device = device
        self.K = K
        self.sensor_size = sensor_size
        self.fov_x, self.fov_y = self._intrinsic_to_fov()
        self.P = self._get_projection_matrix()
        self._pose_cache = (PoseSettingsCache(self._projection_host(), device) if fast_camera
                            else None)

    def get_gaussian_rasterizer(self, cam_position, cam_quaternion):
        return GaussianRasterizer(
            raster_settings=self._get_gaussian_rasterization_settings(cam_position, cam_quaternion))

    def forward(self, points, cam_position=None, cam_quaternion=None, gaussian_rasterizer=None):
        _, M = points.shape
        assert M == 14, "The input tensor should have 14 channels."
        if gaussian_rasterizer is None:
            gaussian_rasterizer = self.get_gaussian_rasterizer(cam_position, cam_quaternion)
        return self._get_gaussian_rasterization(points, gaussian_rasterizer)

    def _intrinsic_to_fov(self):
        fx, fy = self.K[0, 0], self.K[1, 1]
        return (2 * np.arctan2(self.sensor_size[0], 2 * fx),
                2 * np.arctan2(self.sensor_size[1], 2 * fy))

    def _get_projection_matrix(self):
        return torch.from_numpy(self._projection_host()).to(self.device)

    def _projection_host(self):
        fx, fy, cx, cy = self.K[0, 0], self.K[1, 1], self.K[0, 2], self.K[1, 2]
        w, h = self.sensor_size[0], self.sensor_size[1]
        zn, zf = self.z_near, self.z_far
        P = np.zeros((4, 4), dtype=np.float32)
        P[0, 0] = 2.0 * fx / w
        P[1, 1] = 2.0 * fy / h
        P[0, 2] = (2.0 * cx / w) - 1.0
        P[1, 2] = (2.0 * cy / h) - 1.0
        P[2, 2] = -(zf + zn) / (zf - zn)
        P[3, 2] = -1.0
        P[2, 3] = -2.0 * zf * zn / (zf - zn)
        return P

    def _get_w2c_matrix(self, cam_position, cam_quaternion):
        return torch.from_numpy(w2c_host(cam_position, cam_quaternion)).to(self.device)

    def _get_gaussian_rasterization_settings(self, cam_position, cam_quaternion):
        if self._pose_cache is not None:
            view, proj, campos, bg = self._pose_cache.get(cam_position, cam_quaternion)
            return GaussianRasterizationSettings(
                img_h=self.sensor_size[1], img_w=self.sensor_size[0],
                tanfovx=math.tan(self.fov_x * 0.5), tanfovy=math.tan(self.fov_y * 0.5), bg=bg,
                scale_modifier=1.0, view_matrix=view, proj_matrix=proj, sh_degree=0, campos=campos,
                prefiltered=False, debug=False)
        bg = torch.tensor([0.0, 0.0, 0.0], dtype=torch.float32, device=self.device)
        w2c = self._get_w2c_matrix(cam_position, cam_quaternion).transpose(0, 1)
        return GaussianRasterizationSettings(
            img_h=self.sensor_size[1],
            img_w=self.sensor_size[0],
            tanfovx=math.tan(self.fov_x * 0.5),
            tanfovy=math.tan(self.fov_y * 0.5),
            bg=bg,
            scale_modifier=1.0,
            view_matrix=w2c,
            proj_matrix=w2c @ self.P.transpose(0, 1),
            sh_degree=0,
            campos=w2c.inverse()[3, :3],
            prefiltered=False,
            debug=False,
        )

    def _get_gaussian_rasterization(self, points, rasterizer):
        xyz, opacity = points[:, 0:3], points[:, 3:4]
        scales, quaternion, rgbs = points[:, 4:7], points[:, 7:11], points[:, 11:]
        rendered_image, _ = rasterizer(
            means3D=xyz,
            means2D=torch.zeros_like(xyz, dtype=torch.float32, device=self.device),
            shs=None,
            colors_precomp=rgbs,
            opacities=opacity,
            scales=scales,
            rotations=quaternion,
            cov3D_precomp=None,
        )
        if self.flip_lr:
            rendered_image = torch.flip(rendered_image, dims=[2])
        if self.flip_ud:
            rendered_image = torch.flip(rendered_image, dims=[1])
        return rendered_image


def mark_visible(means3D, view_matrix, proj_matrix):
    return dgr_ext.mark_visible(means3D, view_matrix, proj_matrix)
