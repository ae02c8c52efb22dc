"""Forward-only frame loop (scripts/inference.py:208-274, 655-667).

The reference renders a video one frame at a time: per frame it builds the [1,N,14] points,
calls the training wrapper under torch.no_grad() (which still allocates the three opaque buffers
kept for a backward that never comes, zero-fills the image, builds the camera settings with
scipy + four small host->device copies), converts the image on the host and writes it out,
all on one stream.  Here:

  * `FrameRenderer.render(...)` is a forward-only call: no autograd graph, and the geometry /
    binning / image buffers are persistent and only ever grow (the allocator callback hands the
    same storage back), so a steady-state frame allocates nothing;
  * camera settings come from the wrapper's packed host-side path (one upload per new pose, LRU
    cached: gaussiancity_b200/camera.py);
  * attributes go in as structure-of-arrays tensors, opacity / rotation may be omitted
    (adapter.py), an optional crop is a pixel window inside the rasterizer;
  * `render_video(...)` converts each frame to uint8 HWC on the device and copies it into one of
    two pinned host buffers on a side stream, so the device->host copy of frame i overlaps the
    rendering of frame i+1; the caller's `sink(index, ndarray)` runs while the GPU is busy.
"""
import ctypes

import numpy as np
import torch

from . import _cabi, ext
from .adapter import crop_to_window

__all__ = ["FrameRenderer"]


class _GrowingBuffer:
    """Allocator callback target whose storage persists across frames (forward-only: nothing
    has to outlive the call)."""

    def __init__(self, device):
        self.device = device
        self.tensor = torch.empty(0, dtype=torch.uint8, device=device)
        box = self

        def _alloc(_ctx, nbytes):
            try:
                if box.tensor.numel() < nbytes:
                    box.tensor = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, device=device)
                return box.tensor.data_ptr()
            except Exception:
                return 0
        self.cb = _cabi.ALLOC_FN(_alloc)


class FrameRenderer:
    def __init__(self, wrapper):
        """wrapper: gaussiancity_b200.GaussianRasterizerWrapper (camera intrinsics, flips, device);
        build it with fast_camera=True to take the host-side settings path."""
        self.wrapper = wrapper
        self.device = torch.device(wrapper.device)
        self._geom = _GrowingBuffer(self.device)
        self._bin = _GrowingBuffer(self.device)
        self._img = _GrowingBuffer(self.device)
        self._radii = torch.empty(0, dtype=torch.int32, device=self.device)
        self._copy_stream = None
        self.last_num_rendered = 0

    @torch.no_grad()
    def render(self, xyz, scales, rgb, cam_position, cam_quaternion, opacity=None, rotations=None, crop=None,
               out=None):
        """-> image [3,h,w] fp32, equal to wrapper(points14, cam_position, cam_quaternion) (cropped)."""
        w_ = self.wrapper
        rs = w_._get_gaussian_rasterization_settings(cam_position, cam_quaternion)
        W, H = int(rs.img_w), int(rs.img_h)
        x, y, w, h = crop_to_window(W, H, crop, w_.flip_lr, w_.flip_ud)
        dev = self.device
        P = int(xyz.shape[0])
        lib = _cabi.lib()
        with torch.cuda.device(dev):
            if out is None or tuple(out.shape) != (3, h, w):
                out = torch.empty((3, h, w), dtype=torch.float32, device=dev)
            if P == 0:
                return out.zero_()
            if self._radii.numel() < P:
                self._radii = torch.empty(int(P * 1.25) + 1024, dtype=torch.int32, device=dev)
            p = lambda t, name, align=4: ext._ptr(ext._prep(t, name, dev, align))
            keep = [ext._prep(t, n, dev, a) for t, n, a in ((xyz, "means3D", 4), (rgb, "colors_precomp", 4),
                                                            (scales, "scales", 4), (opacity, "opacity", 4),
                                                            (rotations, "rotations", 16))]
            rc = lib.gcr_rasterizer_forward_window(
                self._geom.cb, None, self._bin.cb, None, self._img.cb, None, P, 0, 0, ext._ptr(rs.bg), W, H,
                ext._ptr(keep[0]), None, ext._ptr(keep[1]), ext._ptr(keep[3]), ext._ptr(keep[2]),
                float(rs.scale_modifier), ext._ptr(keep[4]), None, ext._ptr(rs.view_matrix), ext._ptr(rs.proj_matrix),
                None, float(rs.tanfovx), float(rs.tanfovy), 0, ext._ptr(out), ctypes.c_void_p(self._radii.data_ptr()),
                0, x, y, w, h, ext._stream_ptr(dev))
            self.last_num_rendered = _cabi.check(rc, "FrameRenderer.render")
        if w_.flip_lr:
            out = torch.flip(out, dims=[2])
        if w_.flip_ud:
            out = torch.flip(out, dims=[1])
        return out

    @torch.no_grad()
    def render_video(self, frames, sink, crop=None):
        """frames: iterable of dicts(xyz=, scales=, rgb=, cam_position=, cam_quaternion=[, opacity=,
        rotations=]).  Each frame is rendered, mapped like utils.helpers.tensor_to_image
        ((img / 2 + 0.5) clamped, HWC) to uint8 on the device, and handed to sink(index, ndarray[h,w,3])
        -- the copy of frame i to pinned host memory and its sink call overlap the rendering of
        frame i+1.  Returns the number of frames."""
        dev = self.device
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        host, done = [None, None], [None, None]
        pending = []   # (index, slot)
        n = 0

        def drain(upto):
            while len(pending) > upto:
                idx, slot = pending.pop(0)
                done[slot].synchronize()
                sink(idx, host[slot].numpy())

        for i, f in enumerate(frames):
            img = self.render(f["xyz"], f["scales"], f["rgb"], f["cam_position"], f["cam_quaternion"],
                              opacity=f.get("opacity"), rotations=f.get("rotations"), crop=crop)
            u8 = ((img * 0.5 + 0.5).clamp_(0, 1) * 255.0).round_().to(torch.uint8).permute(1, 2, 0).contiguous()
            slot = i & 1
            drain(1)   # the slot written two frames ago must have been consumed
            if host[slot] is None or host[slot].shape != u8.shape:
                host[slot] = torch.empty(u8.shape, dtype=torch.uint8).pin_memory()
                done[slot] = torch.cuda.Event()
            self._copy_stream.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(self._copy_stream):
                host[slot].copy_(u8, non_blocking=True)
                done[slot].record(self._copy_stream)
            u8.record_stream(self._copy_stream)
            pending.append((i, slot))
            n += 1
        drain(0)
        return n
