"""ctypes front-end of the CPU oracle of the hash-grid encoder (oracle/grid_oracle.c) plus an
independent fp64 PyTorch restatement whose autograd is the gradient ground truth.

TEST INFRASTRUCTURE -- only tests/, __graft_entry__ and bench_grid_encoder.py's cpu_baseline
leg may import this module; gaussiancity_b200 never does.
"""
import ctypes
import math
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "grid_oracle.c")
BUILD = os.path.join(HERE, "_build")
SO = os.path.join(BUILD, "libgrid_oracle.so")

_lib = None


def build(force=False):
    os.makedirs(BUILD, exist_ok=True)
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
                               SRC, "-o", SO, "-lm"])
    return SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.ggo_level_scale.restype = ctypes.c_float
        _lib.ggo_level_scale.argtypes = [ctypes.c_uint32, ctypes.c_float, ctypes.c_uint32]
        for n in ("ggo_forward", "ggo_backward_grid", "ggo_backward_input", "ggo_corner_rows"):
            getattr(_lib, n).restype = None
    return _lib


def level_offsets(in_channels, n_levels, base_resolution=16, per_level_scale=2, log2_hashmap_size=19,
                  align_corners=False):
    """The level table sizes of GridEncoder.__init__ (extensions/grid_encoder/__init__.py:140-153).
    Note the reference sizes the tables with the CONSTRUCTOR's per_level_scale argument (default 2),
    not with self.per_level_scale derived from desired_resolution -- restated as is."""
    offsets, offset = [], 0
    max_params = 2 ** log2_hashmap_size
    for i in range(n_levels):
        resolution = int(math.ceil(base_resolution * per_level_scale ** i))
        n = min(max_params, (resolution if align_corners else resolution + 1) ** in_channels)
        n = int(math.ceil(n / 8) * 8)
        offsets.append(offset)
        offset += n
    offsets.append(offset)
    return np.array(offsets, dtype=np.int32)


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _scales(scales, L):
    if scales is None:
        return None, None
    a = _f32(scales).reshape(-1)
    assert a.shape[0] == L
    return a, _p(a)


def level_scale(level, per_level_scale, base_resolution, exp2_ulps=0):
    """fma(exp2f(level * S), H, -1) as the oracle evaluates it; `exp2_ulps` moves the exp2f result by
    that many float steps (the GPU's MUFU.EX2 is within 2 ulp of libm's correctly rounded value)."""
    S = np.float32(math.log2(per_level_scale))
    e = np.float32(np.exp2(np.float32(level) * S))
    for _ in range(abs(exp2_ulps)):
        e = np.nextafter(e, np.float32(np.inf if exp2_ulps > 0 else -np.inf), dtype=np.float32)
    return np.float32(float(e) * float(base_resolution) - 1.0)


def forward(inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False, gridtype=0,
            align_corners=False, scales=None):
    """kernel_grid.  Returns (outputs [L,B,C], dy_dx [B,L*D*C] or None).  `scales`: optional per-level
    override of the level scale (see grid_oracle.c)."""
    inputs, embeddings = _f32(inputs), _f32(embeddings)
    offsets = np.ascontiguousarray(offsets, dtype=np.int32)
    B, D = inputs.shape
    C, L = embeddings.shape[1], offsets.shape[0] - 1
    S = math.log2(per_level_scale)
    out = np.empty((L, B, C), np.float32)
    dy_dx = np.empty((B, L * D * C), np.float32) if calc_grad_inputs else None
    sc, sc_p = _scales(scales, L)
    lib().ggo_forward(_p(inputs), _p(embeddings), _p(offsets), _p(out), ctypes.c_uint32(B), ctypes.c_uint32(D),
                      ctypes.c_uint32(C), ctypes.c_uint32(L), ctypes.c_float(S), ctypes.c_uint32(base_resolution),
                      ctypes.c_int(1 if calc_grad_inputs else 0), _p(dy_dx) if calc_grad_inputs else None,
                      ctypes.c_uint32(gridtype), ctypes.c_int(1 if align_corners else 0), sc_p)
    return out, dy_dx


def backward(grad, inputs, n_embeddings, offsets, per_level_scale, base_resolution, dy_dx=None, gridtype=0,
             align_corners=False, scales=None):
    """kernel_grid_backward (+ kernel_input_backward when dy_dx is given).  grad is [L,B,C].
    Returns (grad_embeddings [sO,C], grad_inputs [B,D] or None)."""
    grad, inputs = _f32(grad), _f32(inputs)
    offsets = np.ascontiguousarray(offsets, dtype=np.int32)
    L, B, C = grad.shape
    D = inputs.shape[1]
    S = math.log2(per_level_scale)
    ge = np.zeros((n_embeddings, C), np.float32)
    sc, sc_p = _scales(scales, L)
    lib().ggo_backward_grid(_p(grad), _p(inputs), _p(offsets), _p(ge), ctypes.c_uint32(B), ctypes.c_uint32(D),
                            ctypes.c_uint32(C), ctypes.c_uint32(L), ctypes.c_float(S),
                            ctypes.c_uint32(base_resolution), ctypes.c_uint32(gridtype),
                            ctypes.c_int(1 if align_corners else 0), sc_p)
    gi = None
    if dy_dx is not None:
        dy_dx = _f32(dy_dx)
        gi = np.empty((B, D), np.float32)
        lib().ggo_backward_input(_p(grad), _p(dy_dx), _p(gi), ctypes.c_uint32(B), ctypes.c_uint32(D),
                                 ctypes.c_uint32(C), ctypes.c_uint32(L))
    return ge, gi


def corner_rows(inputs, offsets, per_level_scale, base_resolution, gridtype=0, align_corners=False, scales=None):
    """Integer work only: table row of each of the 2^D corners, [L,B,2^D] uint32, and their weights."""
    inputs = _f32(inputs)
    offsets = np.ascontiguousarray(offsets, dtype=np.int32)
    B, D = inputs.shape
    L = offsets.shape[0] - 1
    rows = np.empty((L, B, 1 << D), np.uint32)
    w = np.empty((L, B, 1 << D), np.float32)
    sc, sc_p = _scales(scales, L)
    lib().ggo_corner_rows(_p(inputs), _p(offsets), _p(rows), _p(w), ctypes.c_uint32(B), ctypes.c_uint32(D),
                          ctypes.c_uint32(L), ctypes.c_float(math.log2(per_level_scale)),
                          ctypes.c_uint32(base_resolution), ctypes.c_uint32(gridtype),
                          ctypes.c_int(1 if align_corners else 0), sc_p)
    return rows, w


# ---- independent restatement: vectorised PyTorch, any float dtype, differentiable -----------------
_PRIMES = (1, 2654435761, 805459861, 3674653429, 2097192037, 1434869437, 2165219737)


def torch_encode(inputs, embeddings, offsets, per_level_scale, base_resolution, gridtype=0, align_corners=False):
    """outputs [B, L*C] exactly as GridEncoderFunction returns them (extensions/grid_encoder/__init__.py:77),
    written with index arithmetic in int64 (mod 2^32 where the device wraps) and differentiable w.r.t.
    both `inputs` and `embeddings` -- torch.autograd through this in fp64 is the gradient ground truth.
    The level scale is evaluated in fp32 like the device does (libm exp2f), then promoted."""
    import torch
    B, D = inputs.shape
    C = embeddings.shape[1]
    offs = [int(o) for o in offsets]
    L = len(offs) - 1
    S = np.float32(math.log2(per_level_scale))
    outs = []
    M32 = (1 << 32) - 1
    for level in range(L):
        hs = offs[level + 1] - offs[level]
        # fma(exp2f(level * S), H, -1): the product of two floats is exact in fp64, one rounding at the end
        scale32 = np.float32(float(np.float32(np.exp2(np.float32(level) * S))) * float(base_resolution) - 1.0)
        resolution = int(math.ceil(float(scale32))) + 1
        step = resolution if align_corners else resolution + 1
        scale = float(scale32)
        pos = inputs * scale + (0.0 if align_corners else 0.5)
        # the device rounds the position to fp32 once (one FMA) before splitting it into cell and
        # fraction; at resolution ~2000 that rounding is 6e-5 of a cell.  Reproduce the rounded value
        # as a constant and keep the exact expression for the derivative (d frac / d x = scale).
        pos32 = pos.detach().double().float().to(pos.dtype)
        cell = torch.floor(pos32).to(torch.int64)
        frac = (pos32 - cell.to(pos.dtype)) + (pos - pos.detach())
        inside = ((inputs.detach() >= 0) & (inputs.detach() <= 1)).all(dim=1)
        # which leading dimensions the dense index covers (stride <= table size at loop entry)
        strides, stride = [], 1
        for d in range(D):
            if stride <= hs:
                strides.append(stride)
                stride = (stride * step) & M32
            else:
                strides.append(0)
        hashed = gridtype == 0 and stride > hs
        acc = torch.zeros(B, C, dtype=embeddings.dtype, device=embeddings.device)
        for corner in range(1 << D):
            w = torch.ones(B, dtype=inputs.dtype, device=inputs.device)
            idx = torch.zeros(B, dtype=torch.int64, device=inputs.device)
            for d in range(D):
                up = (corner >> d) & 1
                w = w * (frac[:, d] if up else 1 - frac[:, d])
                g = cell[:, d] + up
                if hashed:
                    idx = idx ^ ((g * _PRIMES[d]) & M32)
                else:
                    idx = (idx + g * strides[d]) & M32
            idx = idx % hs
            acc = acc + w[:, None].to(embeddings.dtype) * embeddings[offs[level] + idx]
        outs.append(torch.where(inside[:, None], acc, torch.zeros_like(acc)))
    return torch.stack(outs, dim=1).reshape(B, L * C)
