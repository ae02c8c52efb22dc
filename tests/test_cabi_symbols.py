"""CPU: the C-ABI library loads and exports every symbol include/gcr_rasterizer.h declares
(no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gcr_rasterizer.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"GCR_API\s+[\w\s\*]+?\b(gcr_\w+)\s*\(", src)))


def test_header_declares_the_three_reference_entry_points():
    syms = declared_symbols()
    for s in ("gcr_rasterizer_forward", "gcr_rasterizer_backward", "gcr_rasterizer_mark_visible"):
        assert s in syms


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"not exported: {missing}"
    lib.gcr_abi_version.restype = ctypes.c_int
    assert lib.gcr_abi_version() == 1


def test_python_binding_lists_the_same_symbols(built_lib):
    from gaussiancity_b200 import _cabi
    assert sorted(_cabi.EXPORTED_SYMBOLS) == declared_symbols()
    assert _cabi.lib().gcr_abi_version() == 1


def test_debug_offsets_are_256_aligned_and_ordered(built_lib):
    from gaussiancity_b200 import _cabi
    l = _cabi.lib()
    P, R, W, H = 1000, 5000, 130, 70
    geom = [l.gcr_debug_offset(w, P, R, W, H) for w in range(8)]
    assert all(o % 256 == 0 for o in geom)
    assert l.gcr_debug_offset(_cabi.GEOM_TOTAL_BYTES, P, R, W, H) > max(geom)
    assert l.gcr_debug_offset(_cabi.BIN_TOTAL_BYTES, P, R, W, H) >= 64 * R
    assert l.gcr_debug_offset(12345, P, R, W, H) == ctypes.c_size_t(-1).value


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from gaussiancity_b200 import _cabi
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_cabi.GcrLibraryError, match="no CPU fallback"):
        _cabi.lib()


def test_argument_validation_returns_errors_without_touching_the_device(built_lib):
    """Bad arguments are rejected before any CUDA call: safe to exercise on a GPU-less box."""
    import ctypes as C
    from gaussiancity_b200 import _cabi
    l = _cabi.lib()
    cb = _cabi.ALLOC_FN(lambda ctx, n: 0)
    P = C.c_void_p
    good = C.c_void_p(0x10000)   # never dereferenced on the host

    def fwd(**kw):
        a = dict(geom=cb, binning=cb, img=cb, P=4, D=0, M=1, bg=good, W=16, H=16, means=good, shs=good,
                 colors=None, opac=good, scales=good, rot=good, cov=None, view=good, proj=good, campos=good,
                 out=good, radii=None, shard_rank=0, shard_count=1)
        a.update(kw)
        return l.gcr_rasterizer_forward(a["geom"], None, a["binning"], None, a["img"], None, a["P"], a["D"], a["M"],
                                        a["bg"], a["W"], a["H"], a["means"], a["shs"], a["colors"], a["opac"],
                                        a["scales"], 1.0, a["rot"], a["cov"], a["view"], a["proj"], a["campos"],
                                        1.0, 1.0, 0, a["out"], a["radii"], 0, a["shard_rank"], a["shard_count"], None)

    assert fwd(P=0) == 0                                     # empty input: nothing to do
    cases = [(dict(W=0), "image size"), (dict(shard_rank=2, shard_count=2), "shard"),
             (dict(shs=None, colors=None), "SHs or precomputed"), (dict(scales=None), "scale/rotation"),
             (dict(D=3, M=4), "SH degree"), (dict(rot=C.c_void_p(0x10004)), "16-byte aligned"),
             (dict(M=16, D=3, shs=C.c_void_p(0x10010)), "32-byte aligned"), (dict(means=None), "must not be NULL")]
    for kw, msg in cases:
        assert fwd(**kw) < 0, kw
        assert msg in _cabi.last_error(), (kw, _cabi.last_error())
