"""Seam A: the three native entry points of the reference's pybind module
`diff_gaussian_rasterization_ext` (DGR/bindings.cpp:15-18), with the same positional signatures
and return tuples as RasterizeGaussiansCUDA / RasterizeGaussiansBackwardCUDA / markVisible
(DGR/rasterize_points.cu:35-173), implemented over the C ABI (include/gcr_rasterizer.h).

Host-side responsibilities kept from the reference boundary: output allocation, the three
resizable byte buffers (here: torch uint8 tensors handed out through the allocator callback),
`.contiguous()` on every input, empty tensor == absent optional, P == 0 short-circuit.
Differences: kernels run on torch's CURRENT stream (the reference uses the legacy default
stream); gradient tensors are torch.empty (the kernels write every element).
"""
import ctypes

import torch

from . import _cabi

NUM_CHANNELS = 3  # DGR/cuda_rasterizer/config.h:15


def _ptr(t):
    """Device pointer or NULL for an empty tensor (reference: empty tensor -> nullptr)."""
    if t is None or t.numel() == 0:
        return None
    return ctypes.c_void_p(t.data_ptr())


def _prep(t, name, device, align=4):
    if t is None:
        return None
    if t.numel() == 0:
        return t
    if t.device != device:
        raise RuntimeError(f"{name} must be on {device}, got {t.device}")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32, got {t.dtype}")
    t = t.contiguous()
    if t.data_ptr() % align != 0:
        t = t.clone()
    return t


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class _Allocator:
    """Resizable byte buffer handed to the library (resizeFunctional, rasterize_points.cu:27-33)."""

    def __init__(self, device):
        self.tensor = torch.empty(0, dtype=torch.uint8, device=device)
        box = [self.tensor]   # the C callback closes over `box`, not over self: no reference
        self._box = box       # cycle, so the (GB-sized) buffers die by refcount, not by GC

        def _alloc(_ctx, nbytes):
            try:
                box[0] = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
                return box[0].data_ptr()
            except Exception:  # surfaces as "allocation failed" from the library
                return 0
        self.cb = _cabi.ALLOC_FN(_alloc)

    def take(self):
        """The buffer the library asked for (or the empty tensor); drops the callback."""
        t = self._box[0]
        self.cb = None
        self._box = None
        return t


def rasterize_gaussians(background, means3D, colors, opacity, scales, rotations, scale_modifier,
                        cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height,
                        image_width, sh, degree, campos, prefiltered, debug,
                        shard_rank=0, shard_count=1, stripe_bounds=None, balanced=False):
    """-> (num_rendered, color[3,H,W], radii[P] int32, geomBuffer, binningBuffer, imgBuffer)

    shard_count > 1 renders one contiguous tile-row stripe of the frame (the rest of `color` is
    zero); stripe_bounds: int32 CUDA tensor [shard_count + 1] (e.g. from stripe_partition); None =
    equal-height stripes, or, with balanced=True, stripes of about equal tile-instance count cut on
    the device during the projection pass (read them back with stripe_bounds_of)."""
    if means3D.dim() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    if not means3D.is_cuda:
        raise RuntimeError("means3D must be a CUDA tensor (there is no CPU path)")
    lib = _cabi.lib()
    device = means3D.device
    P, H, W = int(means3D.size(0)), int(image_height), int(image_width)

    with torch.cuda.device(device):
        # every pixel is written by the blend when all tile rows are rendered; a tile-row shard
        # leaves the other ranks' rows untouched, so those must start at zero
        out_alloc = torch.empty if (int(shard_count) == 1 and P != 0) else torch.zeros
        out_color = out_alloc((NUM_CHANNELS, H, W), dtype=torch.float32, device=device)
        radii = torch.zeros((P,), dtype=torch.int32, device=device)
        geom, binning, img = _Allocator(device), _Allocator(device), _Allocator(device)
        rendered = 0
        if P != 0:
            M = int(sh.size(1)) if sh.numel() != 0 else 0
            background = _prep(background, "background", device)
            means3D = _prep(means3D, "means3D", device)
            colors = _prep(colors, "colors_precomp", device)
            opacity = _prep(opacity, "opacity", device)
            scales = _prep(scales, "scales", device)
            rotations = _prep(rotations, "rotations", device, align=16)
            cov3D_precomp = _prep(cov3D_precomp, "cov3D_precomp", device)
            viewmatrix = _prep(viewmatrix, "viewmatrix", device)
            projmatrix = _prep(projmatrix, "projmatrix", device)
            sh = _prep(sh, "sh", device, align=32)
            campos = _prep(campos, "campos", device)
            if stripe_bounds is not None:
                if (stripe_bounds.dtype != torch.int32 or stripe_bounds.device != device or
                        stripe_bounds.numel() != int(shard_count) + 1):
                    raise RuntimeError("stripe_bounds must be an int32 tensor [shard_count + 1] on the device")
                stripe_bounds = stripe_bounds.contiguous()
            rc = lib.gcr_rasterizer_forward_striped(
                geom.cb, None, binning.cb, None, img.cb, None,
                P, int(degree), M, _ptr(background), W, H,
                _ptr(means3D), _ptr(sh), _ptr(colors), _ptr(opacity), _ptr(scales),
                float(scale_modifier), _ptr(rotations), _ptr(cov3D_precomp), _ptr(viewmatrix),
                _ptr(projmatrix), _ptr(campos), float(tan_fovx), float(tan_fovy),
                int(bool(prefiltered)), _ptr(out_color), _ptr(radii), int(bool(debug)),
                int(shard_rank), int(shard_count), _ptr(stripe_bounds), int(bool(balanced)),
                _stream_ptr(device))
            rendered = _cabi.check(rc, "rasterize_gaussians")
    return rendered, out_color, radii, geom.take(), binning.take(), img.take()


def rasterize_gaussians_backward(background, means3D, radii, colors, scales, rotations,
                                 scale_modifier, cov3D_precomp, viewmatrix, projmatrix, tan_fovx,
                                 tan_fovy, dL_dout_color, sh, degree, campos, geomBuffer, R,
                                 binningBuffer, imageBuffer, debug):
    """-> (dL_dmeans2D[P,3], dL_dcolors[P,3], dL_dopacity[P,1], dL_dmeans3D[P,3],
           dL_dcov3D[P,6], dL_dsh[P,M,3], dL_dscales[P,3], dL_drotations[P,4])
    (a striped frame goes through rasterize_gaussians_backward_blend / _geometry)"""
    lib = _cabi.lib()
    device = means3D.device
    P = int(means3D.size(0))
    H, W = int(dL_dout_color.size(1)), int(dL_dout_color.size(2))
    M = int(sh.size(1)) if sh.numel() != 0 else 0
    opts = dict(dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        alloc = torch.empty if P != 0 else torch.zeros
        dL_dmeans3D = alloc((P, 3), **opts)
        dL_dmeans2D = alloc((P, 3), **opts)
        dL_dcolors = alloc((P, NUM_CHANNELS), **opts)
        dL_dopacity = alloc((P, 1), **opts)
        dL_dcov3D = alloc((P, 6), **opts)
        dL_dsh = alloc((P, M, 3), **opts)
        dL_dscales = alloc((P, 3), **opts)
        dL_drotations = alloc((P, 4), **opts)
        if P != 0:
            background = _prep(background, "background", device)
            means3D = _prep(means3D, "means3D", device)
            colors = _prep(colors, "colors_precomp", device)
            scales = _prep(scales, "scales", device)
            rotations = _prep(rotations, "rotations", device, align=16)
            cov3D_precomp = _prep(cov3D_precomp, "cov3D_precomp", device)
            viewmatrix = _prep(viewmatrix, "viewmatrix", device)
            projmatrix = _prep(projmatrix, "projmatrix", device)
            sh = _prep(sh, "sh", device, align=32)
            campos = _prep(campos, "campos", device)
            dL_dout_color = _prep(dL_dout_color, "dL_dout_color", device)
            radii = radii.contiguous()
            rc = lib.gcr_rasterizer_backward(
                P, int(degree), M, int(R), _ptr(background), W, H, _ptr(means3D), _ptr(sh),
                _ptr(colors), _ptr(scales), float(scale_modifier), _ptr(rotations),
                _ptr(cov3D_precomp), _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos),
                float(tan_fovx), float(tan_fovy), _ptr(radii),
                ctypes.c_void_p(geomBuffer.data_ptr()),
                ctypes.c_void_p(binningBuffer.data_ptr()) if binningBuffer.numel() else None,
                ctypes.c_void_p(imageBuffer.data_ptr()),
                _ptr(dL_dout_color), _ptr(dL_dmeans2D), None, _ptr(dL_dopacity),
                _ptr(dL_dcolors), _ptr(dL_dmeans3D), _ptr(dL_dcov3D), _ptr(dL_dsh),
                _ptr(dL_dscales), _ptr(dL_drotations), int(bool(debug)),
                0, 1, _stream_ptr(device))
            _cabi.check(rc, "rasterize_gaussians_backward")
    return (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales,
            dL_drotations)


def mark_visible(means3D, viewmatrix, projmatrix):
    """-> bool[P]: view-space depth > 0.2 (checkFrustum, rasterizer_impl.cu:52-62)"""
    lib = _cabi.lib()
    device = means3D.device
    if not means3D.is_cuda:
        raise RuntimeError("means3D must be a CUDA tensor (there is no CPU path)")
    P = int(means3D.size(0))
    with torch.cuda.device(device):
        present = torch.zeros((P,), dtype=torch.bool, device=device)
        if P != 0:
            means3D = _prep(means3D, "means3D", device)
            viewmatrix = _prep(viewmatrix, "viewmatrix", device)
            projmatrix = _prep(projmatrix, "projmatrix", device)
            rc = lib.gcr_rasterizer_mark_visible(P, _ptr(means3D), _ptr(viewmatrix),
                                                 _ptr(projmatrix), _ptr(present),
                                                 _stream_ptr(device))
            _cabi.check(rc, "mark_visible")
    return present


# ---- crop folded into the rasterizer + GaussianCity's constant attributes (adapter.py) ----
def rasterize_gaussians_window(background, means3D, colors, opacity, scales, rotations, scale_modifier,
                               viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width,
                               window, debug=False):
    """colors_precomp path of rasterize_gaussians restricted to the pixel window (x, y, w, h) of the
    image_height x image_width frame: -> (num_rendered, color[3,h,w], radii, geom, binning, img).
    opacity=None renders every Gaussian opaque, rotations=None uses the identity rotation."""
    lib = _cabi.lib()
    device = means3D.device
    if not means3D.is_cuda:
        raise RuntimeError("means3D must be a CUDA tensor (there is no CPU path)")
    P, H, W = int(means3D.size(0)), int(image_height), int(image_width)
    x, y, w, h = (int(v) for v in window)
    with torch.cuda.device(device):
        out_color = (torch.empty if P != 0 else torch.zeros)((NUM_CHANNELS, h, w), dtype=torch.float32, device=device)
        radii = torch.zeros((P,), dtype=torch.int32, device=device)
        geom, binning, img = _Allocator(device), _Allocator(device), _Allocator(device)
        rendered = 0
        if P != 0:
            background = _prep(background, "background", device)
            means3D = _prep(means3D, "means3D", device)
            colors = _prep(colors, "colors_precomp", device)
            opacity = _prep(opacity, "opacity", device)
            scales = _prep(scales, "scales", device)
            rotations = _prep(rotations, "rotations", device, align=16)
            viewmatrix = _prep(viewmatrix, "viewmatrix", device)
            projmatrix = _prep(projmatrix, "projmatrix", device)
            rc = lib.gcr_rasterizer_forward_window(
                geom.cb, None, binning.cb, None, img.cb, None, P, 0, 0, _ptr(background), W, H,
                _ptr(means3D), None, _ptr(colors), _ptr(opacity), _ptr(scales), float(scale_modifier),
                _ptr(rotations), None, _ptr(viewmatrix), _ptr(projmatrix), None, float(tan_fovx),
                float(tan_fovy), 0, _ptr(out_color), _ptr(radii), int(bool(debug)), x, y, w, h,
                _stream_ptr(device))
            rendered = _cabi.check(rc, "rasterize_gaussians_window")
    return rendered, out_color, radii, geom.take(), binning.take(), img.take()


def rasterize_gaussians_backward_window(background, means3D, radii, scales, rotations, scale_modifier,
                                        viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width,
                                        window, dL_dout_color, geomBuffer, R, binningBuffer, imageBuffer,
                                        want_opacity=False, debug=False):
    """-> (dL_dmeans3D[P,3], dL_dcolors[P,3], dL_dscales[P,3], dL_dopacity[P,1] or None,
    dL_drotations[P,4] or None, dL_dmeans2D[P,3]); dL_dout_color is [3,h,w]."""
    lib = _cabi.lib()
    device = means3D.device
    P = int(means3D.size(0))
    x, y, w, h = (int(v) for v in window)
    opts = dict(dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        alloc = torch.empty if P != 0 else torch.zeros
        dL_dmeans3D, dL_dmeans2D, dL_dcolors = alloc((P, 3), **opts), alloc((P, 3), **opts), alloc((P, 3), **opts)
        dL_dcov3D, dL_dscales = alloc((P, 6), **opts), alloc((P, 3), **opts)
        dL_dopacity = alloc((P, 1), **opts) if want_opacity else None
        dL_drot = alloc((P, 4), **opts) if rotations is not None and rotations.numel() else None
        if P != 0:
            background = _prep(background, "background", device)
            means3D = _prep(means3D, "means3D", device)
            scales = _prep(scales, "scales", device)
            rotations = _prep(rotations, "rotations", device, align=16)
            viewmatrix = _prep(viewmatrix, "viewmatrix", device)
            projmatrix = _prep(projmatrix, "projmatrix", device)
            dL_dout_color = _prep(dL_dout_color, "dL_dout_color", device)
            rc = lib.gcr_rasterizer_backward_window(
                P, 0, 0, int(R), _ptr(background), int(image_width), int(image_height), _ptr(means3D), None,
                None, _ptr(scales), float(scale_modifier), _ptr(rotations), None, _ptr(viewmatrix),
                _ptr(projmatrix), None, float(tan_fovx), float(tan_fovy), _ptr(radii.contiguous()),
                ctypes.c_void_p(geomBuffer.data_ptr()),
                ctypes.c_void_p(binningBuffer.data_ptr()) if binningBuffer.numel() else None,
                ctypes.c_void_p(imageBuffer.data_ptr()), _ptr(dL_dout_color), _ptr(dL_dmeans2D), None,
                _ptr(dL_dopacity), _ptr(dL_dcolors), _ptr(dL_dmeans3D), _ptr(dL_dcov3D), None,
                _ptr(dL_dscales), _ptr(dL_drot), int(bool(debug)), x, y, w, h, _stream_ptr(device))
            _cabi.check(rc, "rasterize_gaussians_backward_window")
    return dL_dmeans3D, dL_dcolors, dL_dscales, dL_dopacity, dL_drot, dL_dmeans2D


# ---- tile-row stripes of a multi-GPU frame (gaussiancity_b200.sharding) -------------------------
def stripe_partition(means3D, scales, rotations, scale_modifier, viewmatrix, projmatrix, tan_fovx,
                     tan_fovy, image_height, image_width, shard_count, workspace, bounds_out,
                     cov3D_precomp=None):
    """Balanced contiguous tile-row stripes -> bounds_out (int32 [shard_count + 1], device), no
    host synchronisation.  workspace: int32 [ceil(H/16) + 1], device."""
    lib = _cabi.lib()
    device = means3D.device
    P = int(means3D.size(0))
    e = torch.Tensor([])
    with torch.cuda.device(device):
        means3D = _prep(means3D, "means3D", device)
        scales = _prep(scales if scales is not None else e, "scales", device)
        rotations = _prep(rotations if rotations is not None else e, "rotations", device, align=16)
        cov3D_precomp = _prep(cov3D_precomp if cov3D_precomp is not None else e, "cov3D_precomp", device)
        viewmatrix = _prep(viewmatrix, "viewmatrix", device)
        projmatrix = _prep(projmatrix, "projmatrix", device)
        rc = lib.gcr_stripe_partition(P, _ptr(means3D), _ptr(scales), float(scale_modifier), _ptr(rotations),
                                      _ptr(cov3D_precomp), _ptr(viewmatrix), _ptr(projmatrix),
                                      int(image_width), int(image_height), float(tan_fovx), float(tan_fovy),
                                      int(shard_count), _ptr(workspace), _ptr(bounds_out), _stream_ptr(device))
        _cabi.check(rc, "stripe_partition")
    return bounds_out


def stripe_bounds_of(geomBuffer, P, shard_count):
    """int32 [shard_count + 1]: the tile-row stripe bounds the last forward into geomBuffer used."""
    lib = _cabi.lib()
    off = lib.gcr_debug_offset(_cabi.GEOM_COUNTERS, int(P), 0, 16, 16) + 4 * 8
    base = (-geomBuffer.data_ptr()) % 256
    return geomBuffer[base + off:base + off + 4 * (int(shard_count) + 1)].view(torch.int32)


def owner_bytes(geomBuffer, P):
    """uint8 [P]: rank that owns each Gaussian of the last forward into geomBuffer (255 = culled)."""
    lib = _cabi.lib()
    off = lib.gcr_debug_offset(_cabi.GEOM_OWNER, int(P), 0, 16, 16)
    base = (-geomBuffer.data_ptr()) % 256   # the library carves from the 256-byte aligned address
    return geomBuffer[base + off:base + off + int(P)]


def rasterize_gaussians_backward_blend(background, P, R, dL_dout_color, geomBuffer, binningBuffer,
                                       imageBuffer, debug=False, shard_rank=0, shard_count=1,
                                       accumulators=None, remote_scalar=False):
    """First half of the backward: per-tile gradient blend over this rank's stripe.

    accumulators=None -> returns a fresh grad_acc [P,12] fp32 (dmean2D.xy, dconic.xyw, dopacity,
    dcolor.rgb, 3 pad): PARTIAL sums when shard_count > 1 (reduce across ranks before the
    geometry half).  accumulators=[device addresses, one per rank] (peer-mapped, see
    sharding.CudaBackend.peer_setup): each Gaussian's sums are added into its owner's accumulator
    inside the kernel; returns None."""
    lib = _cabi.lib()
    device = dL_dout_color.device
    H, W = int(dL_dout_color.size(1)), int(dL_dout_color.size(2))
    with torch.cuda.device(device):
        grad_acc = None
        if accumulators is None:
            grad_acc = torch.empty((int(P), 12), dtype=torch.float32, device=device)
            ptrs = (ctypes.c_void_p * 1)(grad_acc.data_ptr())
            n_acc, zero_first = 1, 1
        else:
            ptrs = (ctypes.c_void_p * len(accumulators))(*[int(a) for a in accumulators])
            n_acc, zero_first = len(accumulators), 0
        if P != 0:
            background = _prep(background, "background", device)
            dL_dout_color = _prep(dL_dout_color, "dL_dout_color", device)
            rc = lib.gcr_rasterizer_backward_blend(
                int(P), int(R), _ptr(background), W, H, ctypes.c_void_p(geomBuffer.data_ptr()),
                ctypes.c_void_p(binningBuffer.data_ptr()) if binningBuffer.numel() else None,
                ctypes.c_void_p(imageBuffer.data_ptr()), _ptr(dL_dout_color), ptrs, n_acc, zero_first,
                int(bool(remote_scalar)), int(bool(debug)), int(shard_rank), int(shard_count), None,
                _stream_ptr(device))
            _cabi.check(rc, "rasterize_gaussians_backward_blend")
    return grad_acc


def packed_views(rows, dL_dsh):
    """The 8-tuple of rasterize_gaussians_backward as views into a packed [P,24] gradient tensor
    (layout: include/gcr_rasterizer.h, gcr_rasterizer_backward_geometry)."""
    return (rows[:, 7:10], rows[:, 10:13], rows[:, 3:4], rows[:, 0:3], rows[:, 13:19], dL_dsh, rows[:, 4:7],
            rows[:, 20:24])


def rasterize_gaussians_backward_geometry(means3D, radii, scales, rotations, scale_modifier,
                                          cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy,
                                          image_height, image_width, sh, degree, campos, geomBuffer,
                                          grad_acc, range_start=0, range_count=-1, debug=False,
                                          out=None, shard_rank=0, striped=False, clear_accumulator=False,
                                          packed=None):
    """Second half of the backward: per-Gaussian geometry gradients, for the Gaussians within
    [range_start, range_start+range_count) that `shard_rank` owns, from the (reduced) accumulator
    (a [P,12] tensor or a device address).  Returns the same 8-tuple as
    rasterize_gaussians_backward.  striped=False: the rows of all other Gaussians in the range are
    written as zeros; striped=True: they are left untouched (uninitialised unless `out` is
    supplied).  packed=(rows [P,24], dL_dsh [P,M,3]) (or True to allocate them): the kernel writes one
    96-byte row per Gaussian and the returned tensors are views into it (see packed_views)."""
    lib = _cabi.lib()
    device = means3D.device
    P = int(means3D.size(0))
    M = int(sh.size(1)) if sh.numel() != 0 else 0
    opts = dict(dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        rows = None
        if packed is not None and packed is not False:
            if packed is True:
                packed = (torch.empty((P, 24), **opts), torch.empty((P, M, 3), **opts))
            rows, dsh_p = packed
            out = packed_views(rows, dsh_p)
        if out is None:
            out = (torch.empty((P, 3), **opts), torch.empty((P, 3), **opts), torch.empty((P, 1), **opts),
                   torch.empty((P, 3), **opts), torch.empty((P, 6), **opts), torch.empty((P, M, 3), **opts),
                   torch.empty((P, 3), **opts), torch.empty((P, 4), **opts))
        dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drot = out
        if P != 0:
            means3D = _prep(means3D, "means3D", device)
            scales = _prep(scales, "scales", device)
            rotations = _prep(rotations, "rotations", device, align=16)
            cov3D_precomp = _prep(cov3D_precomp, "cov3D_precomp", device)
            viewmatrix = _prep(viewmatrix, "viewmatrix", device)
            projmatrix = _prep(projmatrix, "projmatrix", device)
            sh = _prep(sh, "sh", device, align=32)
            campos = _prep(campos, "campos", device)
            acc_ptr = _ptr(grad_acc) if isinstance(grad_acc, torch.Tensor) else ctypes.c_void_p(int(grad_acc))
            rc = lib.gcr_rasterizer_backward_geometry(
                P, int(degree), M, _ptr(means3D), _ptr(sh), _ptr(scales), float(scale_modifier),
                _ptr(rotations), _ptr(cov3D_precomp), _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos),
                int(image_width), int(image_height), float(tan_fovx), float(tan_fovy),
                _ptr(radii.contiguous()), ctypes.c_void_p(geomBuffer.data_ptr()), acc_ptr,
                None if rows is not None else _ptr(dL_dmeans2D), None,
                None if rows is not None else _ptr(dL_dopacity), None if rows is not None else _ptr(dL_dcolors),
                None if rows is not None else _ptr(dL_dmeans3D), None if rows is not None else _ptr(dL_dcov3D),
                _ptr(dL_dsh), None if rows is not None else _ptr(dL_dscales),
                None if rows is not None else _ptr(dL_drot), int(bool(debug)),
                int(range_start), int(range_count), int(shard_rank), int(bool(striped)),
                int(bool(clear_accumulator)), _ptr(rows) if rows is not None else None, _stream_ptr(device))
            _cabi.check(rc, "rasterize_gaussians_backward_geometry")
    return out
