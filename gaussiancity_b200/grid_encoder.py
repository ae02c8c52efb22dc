"""B200-native multi-resolution hash-grid encoder -- drop-in for hzxie/GaussianCity's
extensions/grid_encoder, the positional encoder of the generator when
POS_EMD = "HASH_GRID" (models/generator.py:35-42, config.py:121-123: 16 levels x 8 channels over
ENCODER_OUT_DIM = 5 coordinates).

Seam B (this module) mirrors extensions/grid_encoder/__init__.py name for name:
  GridEncoderFunction   extensions/grid_encoder/__init__.py:19-124
  GridEncoder           extensions/grid_encoder/__init__.py:127-193
Seam A is the native module `grid_encoder_ext` (gaussiancity_b200/compat, csrc/grid_module.cpp),
or the same two functions over ctypes below; both are thin hosts over the C ABI of
include/gcr_grid_encoder.h (libgcr_grid_encoder.so, csrc/grid_encoder.cu).  There is no CPU path:
a missing library raises, CPU tensors are refused.

Differences from the reference that a caller can observe: embeddings must be float32 (the
reference also dispatches half / double; GaussianCity never creates such a table); the module's
`fused_backward` (default on: both gradients in one launch, no [B, L, D, C] derivative tensor --
measured 0.231 vs 0.267 ms per forward + backward at the generator's size; the input gradient then
agrees with the reference's to 1.5e-7 instead of bit for bit; off = the reference's two-pass
backward through a stored dy_dx tensor, which is also what Seam A always does).
"""
import ctypes
import math
import os

import numpy as np
import torch

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libgcr_grid_encoder.so")
ABI_VERSION = 1   # GCR_GRID_ABI_VERSION

# names must match include/gcr_grid_encoder.h; checked by tests/test_cabi_symbols.py
EXPORTED_SYMBOLS = (
    "gcr_grid_abi_version",
    "gcr_grid_last_error",
    "gcr_grid_encode_forward",
    "gcr_grid_encode_backward",
    "gcr_grid_encode_backward_fused",
)

_lib = None


class GridLibraryError(RuntimeError):
    pass


def lib():
    """The loaded CDLL.  Raises GridLibraryError if the CUDA library is missing (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GridLibraryError(
            f"{LIB_PATH} not found: the sm_100a grid-encoder library has not been built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (or gaussiancity_b200/build.py). "
            "There is no CPU fallback.")
    l = ctypes.CDLL(LIB_PATH)
    for s in EXPORTED_SYMBOLS:
        if not hasattr(l, s):
            raise GridLibraryError(f"{LIB_PATH} does not export {s}")
    u32, f32, vp, ci = ctypes.c_uint32, ctypes.c_float, ctypes.c_void_p, ctypes.c_int
    l.gcr_grid_abi_version.restype = ci
    l.gcr_grid_abi_version.argtypes = []
    l.gcr_grid_last_error.restype = ctypes.c_char_p
    l.gcr_grid_last_error.argtypes = []
    l.gcr_grid_encode_forward.restype = ci
    l.gcr_grid_encode_forward.argtypes = [vp, vp, vp, vp, u32, u32, u32, u32, f32, u32, ci, vp, u32, ci, vp]
    l.gcr_grid_encode_backward.restype = ci
    l.gcr_grid_encode_backward.argtypes = [vp, vp, vp, vp, vp, u32, u32, u32, u32, f32, u32, ci, vp, vp, u32, ci, vp]
    l.gcr_grid_encode_backward_fused.restype = ci
    l.gcr_grid_encode_backward_fused.argtypes = [vp, vp, vp, vp, vp, u32, u32, u32, u32, f32, u32, vp, u32, ci, vp]
    _lib = l
    return l


def _check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what}: {lib().gcr_grid_last_error().decode(errors='replace')}")


def _dev_f32(t, name):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous tensor")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32: the B200 grid encoder implements the fp32 table "
                           "GaussianCity uses")
    return t.data_ptr()


def _dev_i32(t, name):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous tensor")
    if t.dtype != torch.int32:
        raise RuntimeError(f"{name} must be an int tensor")
    return t.data_ptr()


class _CtypesBinding:
    """`grid_encoder_ext.forward / backward` (extensions/grid_encoder/bindings.cpp:35-40) over ctypes."""

    @staticmethod
    def forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, calc_grad_inputs, dy_dx, gridtype,
                align_corners):
        with torch.cuda.device(inputs.device):
            stream = torch.cuda.current_stream().cuda_stream
            _check(lib().gcr_grid_encode_forward(
                _dev_f32(inputs, "inputs"), _dev_f32(embeddings, "embeddings"), _dev_i32(offsets, "offsets"),
                _dev_f32(outputs, "outputs"), B, D, C, L, S, H, 1 if calc_grad_inputs else 0,
                _dev_f32(dy_dx, "dy_dx"), gridtype, 1 if align_corners else 0, stream), "grid_encode_forward")

    @staticmethod
    def backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H, calc_grad_inputs, dy_dx,
                 grad_inputs, gridtype, align_corners):
        with torch.cuda.device(inputs.device):
            stream = torch.cuda.current_stream().cuda_stream
            _check(lib().gcr_grid_encode_backward(
                _dev_f32(grad, "grad"), _dev_f32(inputs, "inputs"), None, _dev_i32(offsets, "offsets"),
                _dev_f32(grad_embeddings, "grad_embeddings"), B, D, C, L, S, H, 1 if calc_grad_inputs else 0,
                _dev_f32(dy_dx, "dy_dx"), _dev_f32(grad_inputs, "grad_inputs"), gridtype,
                1 if align_corners else 0, stream), "grid_encode_backward")

    @staticmethod
    def backward_fused(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H, grad_inputs, gridtype,
                       align_corners):
        with torch.cuda.device(inputs.device):
            stream = torch.cuda.current_stream().cuda_stream
            _check(lib().gcr_grid_encode_backward_fused(
                _dev_f32(grad, "grad"), _dev_f32(inputs, "inputs"), _dev_f32(embeddings, "embeddings"),
                _dev_i32(offsets, "offsets"), _dev_f32(grad_embeddings, "grad_embeddings"), B, D, C, L, S, H,
                _dev_f32(grad_inputs, "grad_inputs"), gridtype, 1 if align_corners else 0, stream),
                "grid_encode_backward_fused")

    @staticmethod
    def abi_version():
        return lib().gcr_grid_abi_version()


ctypes_ext = _CtypesBinding


def _pick_binding():
    want = os.environ.get("GCR_HOST_BINDING", "")
    if want == "ctypes":
        return ctypes_ext, "ctypes"
    try:
        from .compat import grid_encoder_ext as native
        if native.abi_version() == ABI_VERSION:
            return native, "native"
        if want == "native":
            raise ImportError("native grid_encoder_ext was built against another ABI version: rebuild")
    except ImportError:
        if want == "native":
            raise
    return ctypes_ext, "ctypes"


grid_encoder_ext, HOST_BINDING = _pick_binding()


class GridEncoderFunction(torch.autograd.Function):
    """extensions/grid_encoder/__init__.py:19-124, same positional arguments (+ fused_backward)."""

    @staticmethod
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution, *optional):
        # inputs [B, D] in [0, 1]; embeddings [sO, C]; offsets [L + 1] int32; returns [B, L * C]
        # optional, positional like the reference: calc_grad_inputs=False, gridtype=0, align_corners=False
        # (+ fused_backward=False); autograd wants one gradient per argument actually passed
        if len(optional) > 4:
            raise TypeError("GridEncoderFunction takes at most 9 arguments")
        calc_grad_inputs, gridtype, align_corners, fused_backward = \
            tuple(optional) + (False, 0, False, False)[len(optional):]
        ctx.n_optional = len(optional)
        inputs = inputs.contiguous()
        B, D = inputs.shape
        L = offsets.shape[0] - 1
        C = embeddings.shape[1]
        S = math.log2(per_level_scale)
        H = base_resolution
        outputs = torch.empty(L, B, C, device=inputs.device, dtype=embeddings.dtype)
        if calc_grad_inputs and not fused_backward:
            dy_dx = torch.empty(B, L * D * C, device=inputs.device, dtype=embeddings.dtype)
        else:
            dy_dx = torch.empty(1, device=inputs.device, dtype=embeddings.dtype)
        grid_encoder_ext.forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, H,
                                 calc_grad_inputs and not fused_backward, dy_dx, gridtype, align_corners)
        outputs = outputs.permute(1, 0, 2).reshape(B, L * C)
        ctx.save_for_backward(inputs, embeddings, offsets, dy_dx)
        ctx.dims = [B, D, C, L, S, H, gridtype]
        ctx.calc_grad_inputs = calc_grad_inputs
        ctx.align_corners = align_corners
        ctx.fused_backward = fused_backward
        return outputs

    @staticmethod
    def backward(ctx, grad):
        inputs, embeddings, offsets, dy_dx = ctx.saved_tensors
        B, D, C, L, S, H, gridtype = ctx.dims
        calc_grad_inputs = ctx.calc_grad_inputs
        align_corners = ctx.align_corners
        # [B, L * C] -> [L, B, C]
        grad = grad.view(B, L, C).permute(1, 0, 2).contiguous()
        grad_embeddings = torch.zeros_like(embeddings)
        if calc_grad_inputs:
            grad_inputs = torch.zeros_like(inputs, dtype=embeddings.dtype)
        else:
            grad_inputs = torch.zeros(1, device=inputs.device, dtype=embeddings.dtype)
        if calc_grad_inputs and ctx.fused_backward:
            grid_encoder_ext.backward_fused(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H,
                                            grad_inputs, gridtype, align_corners)
        else:
            grid_encoder_ext.backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H,
                                      calc_grad_inputs, dy_dx, grad_inputs, gridtype, align_corners)
        none = (None,) * (3 + ctx.n_optional)
        if calc_grad_inputs:
            grad_inputs = grad_inputs.to(inputs.dtype)
            return (grad_inputs, grad_embeddings) + none
        return (None, grad_embeddings) + none


def level_offsets(in_channels, n_levels, base_resolution=16, per_level_scale=2, log2_hashmap_size=19,
                  align_corners=False):
    """Table offsets of the levels (extensions/grid_encoder/__init__.py:140-153).  As in the reference
    the sizes follow the `per_level_scale` ARGUMENT (default 2), not the scale derived from
    desired_resolution that the kernels use."""
    offsets, offset = [], 0
    max_params = 2 ** log2_hashmap_size
    for i in range(n_levels):
        resolution = int(math.ceil(base_resolution * per_level_scale ** i))
        params_in_level = min(max_params, (resolution if align_corners else resolution + 1) ** in_channels)
        params_in_level = int(math.ceil(params_in_level / 8) * 8)
        offsets.append(offset)
        offset += params_in_level
    offsets.append(offset)
    return np.array(offsets, dtype=np.int32)


class GridEncoder(torch.nn.Module):
    """extensions/grid_encoder/__init__.py:127-193: same constructor, attributes, parameter
    (`embeddings` [sum of level sizes, lvl_channels], U(-1e-4, 1e-4)) and buffer (`offsets`)."""

    def __init__(self, in_channels, n_levels, lvl_channels, desired_resolution, per_level_scale=2,
                 base_resolution=16, log2_hashmap_size=19, gridtype="hash", align_corners=False,
                 fused_backward=True):
        super().__init__()
        self.in_channels = in_channels
        self.n_levels = n_levels
        self.lvl_channels = lvl_channels
        self.per_level_scale = 2 ** (math.log2(desired_resolution / base_resolution) / (n_levels - 1))
        self.log2_hashmap_size = log2_hashmap_size
        self.base_resolution = base_resolution
        self.output_dim = n_levels * lvl_channels
        self.gridtype = gridtype
        self.gridtype_id = 0 if gridtype == "hash" else 1
        self.align_corners = align_corners
        self.fused_backward = fused_backward
        self.max_params = 2 ** log2_hashmap_size
        offsets = level_offsets(in_channels, n_levels, base_resolution, per_level_scale, log2_hashmap_size,
                                align_corners)
        offset = int(offsets[-1])
        offsets = torch.from_numpy(offsets)
        self.register_buffer("offsets", offsets)
        self.n_params = offsets[-1] * lvl_channels
        self.embeddings = torch.nn.Parameter(torch.empty(offset, lvl_channels))
        self._init_weights()

    def _init_weights(self):
        self.embeddings.data.uniform_(-1e-4, 1e-4)

    def forward(self, inputs, bound=1):
        # inputs [..., in_channels] in [-bound, bound] -> [..., n_levels * lvl_channels]
        inputs = (inputs + bound) / (2 * bound)
        prefix_shape = list(inputs.shape[:-1])
        inputs = inputs.view(-1, self.in_channels)
        outputs = GridEncoderFunction.apply(inputs, self.embeddings, self.offsets, self.per_level_scale,
                                            self.base_resolution, inputs.requires_grad, self.gridtype_id,
                                            self.align_corners, self.fused_backward)
        return outputs.view(prefix_shape + [self.output_dim])
