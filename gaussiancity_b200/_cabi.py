"""ctypes binding of the C ABI in include/gcr_rasterizer.h (libgcr_rasterizer.so).

There is NO CPU or PyTorch fallback: if the CUDA library is missing or fails to load, importing
this module's `lib()` raises, loudly.  (`oracle/` is test infrastructure and is never imported
from here.)
"""
import ctypes
import os

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libgcr_rasterizer.so")

ALLOC_FN = ctypes.CFUNCTYPE(ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t)

# names must match include/gcr_rasterizer.h; checked by tests/test_cabi_symbols.py
EXPORTED_SYMBOLS = (
    "gcr_abi_version",
    "gcr_last_error",
    "gcr_rasterizer_forward",
    "gcr_rasterizer_forward_striped",
    "gcr_rasterizer_forward_window",
    "gcr_rasterizer_backward_window",
    "gcr_rasterizer_backward",
    "gcr_rasterizer_backward_blend",
    "gcr_rasterizer_backward_geometry",
    "gcr_rasterizer_mark_visible",
    "gcr_stripe_partition",
    "gcr_peer_alloc",
    "gcr_peer_open",
    "gcr_peer_close",
    "gcr_peer_free",
    "gcr_peer_barrier",
    "gcr_debug_offset",
    "gcr_debug_set_cov3d_out",
    "gcr_profile_enable",
    "gcr_profile_stage_count",
    "gcr_profile_stage_name",
    "gcr_profile_stage_ms",
    "gcr_set_programmatic_launch",
)

# enum values of gcr_debug_offset (include/gcr_rasterizer.h)
GEOM_DEPTH_SORTED_KEYS, GEOM_TILES_TOUCHED, GEOM_RECORDS, GEOM_CLAMPED = 0, 1, 2, 3
GEOM_SORTED_GAUSS, GEOM_OFFSETS, GEOM_GRAD_ACC, GEOM_RADII, GEOM_OWNER, GEOM_COUNTERS = 4, 5, 6, 7, 8, 9
GEOM_TOTAL_BYTES = 100
BIN_POINT_LIST, BIN_TILE_KEYS, BIN_TOTAL_BYTES = 200, 201, 300
ABI_VERSION = 2          # GCR_ABI_VERSION
MAX_SHARDS = 16          # GCR_MAX_SHARDS
PEER_HANDLE_BYTES = 64   # GCR_PEER_HANDLE_BYTES
IMG_FINAL_T, IMG_N_CONTRIB, IMG_RANGES, IMG_TOTAL_BYTES = 400, 401, 402, 500

_lib = None


class GcrLibraryError(RuntimeError):
    pass


def _declare(l):
    c_int, c_float, c_void_p, c_size_t = ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t
    l.gcr_abi_version.restype = c_int
    l.gcr_abi_version.argtypes = []
    l.gcr_last_error.restype = ctypes.c_char_p
    l.gcr_last_error.argtypes = []
    fwd_args = ([ALLOC_FN, c_void_p] * 3 + [c_int] * 3 + [c_void_p, c_int, c_int] +
                [c_void_p] * 5 + [c_float] + [c_void_p] * 5 + [c_float, c_float, c_int] +
                [c_void_p, c_void_p, c_int, c_int, c_int])
    l.gcr_rasterizer_forward.restype = c_int
    l.gcr_rasterizer_forward.argtypes = fwd_args + [c_void_p]
    l.gcr_rasterizer_forward_striped.restype = c_int
    l.gcr_rasterizer_forward_striped.argtypes = fwd_args + [c_void_p, c_int, c_void_p]
    l.gcr_rasterizer_forward_window.restype = c_int
    l.gcr_rasterizer_forward_window.argtypes = fwd_args[:-2] + [c_int] * 4 + [c_void_p]
    bwd_args = ([c_int] * 4 + [c_void_p, c_int, c_int] + [c_void_p] * 4 + [c_float] + [c_void_p] * 5 +
                [c_float, c_float] + [c_void_p] * 14 + [c_int])
    l.gcr_rasterizer_backward.restype = c_int
    l.gcr_rasterizer_backward.argtypes = bwd_args + [c_int, c_int, c_void_p]
    l.gcr_rasterizer_backward_window.restype = c_int
    l.gcr_rasterizer_backward_window.argtypes = bwd_args + [c_int] * 4 + [c_void_p]
    l.gcr_rasterizer_backward_blend.restype = c_int
    l.gcr_rasterizer_backward_blend.argtypes = (
        [c_int, c_int, c_void_p, c_int, c_int] + [c_void_p] * 4 + [ctypes.POINTER(c_void_p)] +
        [c_int] * 6 + [ctypes.POINTER(c_int), c_void_p])
    l.gcr_rasterizer_backward_geometry.restype = c_int
    l.gcr_rasterizer_backward_geometry.argtypes = (
        [c_int] * 3 + [c_void_p] * 3 + [c_float] + [c_void_p] * 5 + [c_int, c_int, c_float, c_float] +
        [c_void_p] * 12 + [c_int] * 6 + [c_void_p, c_void_p])
    l.gcr_stripe_partition.restype = c_int
    l.gcr_stripe_partition.argtypes = ([c_int, c_void_p, c_void_p, c_float] + [c_void_p] * 4 +
                                       [c_int, c_int, c_float, c_float, c_int] + [c_void_p] * 3)
    l.gcr_peer_alloc.restype = c_int
    l.gcr_peer_alloc.argtypes = [c_size_t, ctypes.POINTER(c_void_p), ctypes.POINTER(ctypes.c_ubyte)]
    l.gcr_peer_open.restype = c_int
    l.gcr_peer_open.argtypes = [ctypes.POINTER(ctypes.c_ubyte), ctypes.POINTER(c_void_p)]
    l.gcr_peer_close.restype = c_int
    l.gcr_peer_close.argtypes = [c_void_p]
    l.gcr_peer_free.restype = c_int
    l.gcr_peer_free.argtypes = [c_void_p]
    l.gcr_peer_barrier.restype = c_int
    l.gcr_peer_barrier.argtypes = [ctypes.POINTER(c_void_p), c_int, c_int, ctypes.c_uint, c_void_p]
    l.gcr_rasterizer_mark_visible.restype = c_int
    l.gcr_rasterizer_mark_visible.argtypes = [c_int] + [c_void_p] * 5
    l.gcr_debug_offset.restype = c_size_t
    l.gcr_debug_offset.argtypes = [c_int] * 5
    l.gcr_debug_set_cov3d_out.restype = None
    l.gcr_debug_set_cov3d_out.argtypes = [c_void_p]
    l.gcr_profile_enable.restype = None
    l.gcr_profile_enable.argtypes = [c_int]
    l.gcr_profile_stage_count.restype = c_int
    l.gcr_profile_stage_count.argtypes = []
    l.gcr_profile_stage_name.restype = ctypes.c_char_p
    l.gcr_profile_stage_name.argtypes = [c_int]
    l.gcr_profile_stage_ms.restype = c_float
    l.gcr_profile_stage_ms.argtypes = [c_int]
    l.gcr_set_programmatic_launch.restype = c_int
    l.gcr_set_programmatic_launch.argtypes = [c_int]


def lib():
    """The loaded CDLL. Raises GcrLibraryError if the CUDA extension is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GcrLibraryError(
            f"{LIB_PATH} not found: the sm_100a CUDA library has not been built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (or gaussiancity_b200/build.py). "
            "There is no CPU fallback.")
    try:
        l = ctypes.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise GcrLibraryError(f"failed to load {LIB_PATH}: {e}") from e
    for s in EXPORTED_SYMBOLS:
        if not hasattr(l, s):
            raise GcrLibraryError(f"{LIB_PATH} does not export {s}")
    _declare(l)
    _lib = l
    return l


def last_error():
    return lib().gcr_last_error().decode(errors="replace")


def check(rc, what):
    if rc < 0:
        raise RuntimeError(f"{what}: {last_error()}")
    return rc


def set_programmatic_launch(on):
    """Programmatic dependent launch of the frame's kernel chain (include/gcr_rasterizer.h); returns the
    previous setting."""
    return bool(lib().gcr_set_programmatic_launch(1 if on else 0))


def profile_enable(on=True):
    lib().gcr_profile_enable(1 if on else 0)


def profile_read():
    """{stage name: device ms} of the most recent profiled call (call after synchronising)."""
    l = lib()
    out = {}
    for i in range(l.gcr_profile_stage_count()):
        ms = l.gcr_profile_stage_ms(i)
        if ms >= 0:
            out[l.gcr_profile_stage_name(i).decode()] = float(ms)
    return out
