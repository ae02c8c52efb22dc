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
    assert lib.gcr_abi_version() == 2


def test_python_binding_lists_the_same_symbols(built_lib):
    from gaussiancity_b200 import _cabi
    assert sorted(_cabi.EXPORTED_SYMBOLS) == declared_symbols()
    assert _cabi.lib().gcr_abi_version() == 2


def test_debug_offsets_are_256_aligned_and_ordered(built_lib):
    from gaussiancity_b200 import _cabi
    l = _cabi.lib()
    P, R, W, H = 1000, 5000, 130, 70
    geom = [l.gcr_debug_offset(w, P, R, W, H) for w in range(10)]
    assert all(o % 256 == 0 for o in geom)
    assert l.gcr_debug_offset(_cabi.GEOM_TOTAL_BYTES, P, R, W, H) > max(geom)
    assert l.gcr_debug_offset(_cabi.BIN_TOTAL_BYTES, P, R, W, H) >= 16 * R
    assert l.gcr_debug_offset(_cabi.BIN_POINT_LIST, P, R, W, H) == 0   # the backward relies on it
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
             (dict(M=16, D=3, shs=C.c_void_p(0x10010)), "32-byte aligned"), (dict(means=None), "must not be NULL"),
             (dict(W=70000 * 16), "65520"),
             (dict(M=25, D=3), "at most 16 SH"), (dict(shard_rank=0, shard_count=17), "shard")]
    for kw, msg in cases:
        assert fwd(**kw) < 0, kw
        assert msg in _cabi.last_error(), (kw, _cabi.last_error())

    # backward halves: same discipline (a bad pointer must be an error code, not a device fault)
    def blend(**kw):
        a = dict(P=4, R=8, bg=good, W=16, H=16, geom=good, binning=good, img=good, dpix=good,
                 accs=[0x20000], zero=0, rank=0, count=1)
        a.update(kw)
        arr = (C.c_void_p * max(1, len(a["accs"])))(*a["accs"]) if a["accs"] is not None else None
        return l.gcr_rasterizer_backward_blend(a["P"], a["R"], a["bg"], a["W"], a["H"], a["geom"], a["binning"],
                                               a["img"], a["dpix"], arr, len(a["accs"] or []), a["zero"], 0, 0,
                                               a["rank"], a["count"], None, None)

    assert blend(P=0) == 0
    for kw, msg in [(dict(accs=None), "accumulators"), (dict(accs=[0x20000, 0x30000]), "one per rank"),
                    (dict(accs=[0x20004]), "16-byte aligned"), (dict(accs=[0]), "must not be NULL"),
                    (dict(geom=None), "must not be NULL"), (dict(binning=None), "binning_buffer"),
                    (dict(W=0), "image size"), (dict(rank=1, count=1), "shard")]:
        assert blend(**kw) < 0, kw
        assert msg in _cabi.last_error(), (kw, _cabi.last_error())

    def geometry(**kw):
        a = dict(P=4, D=0, M=1, means=good, shs=good, scales=good, rot=good, cov=None, view=good, proj=good,
                 campos=good, W=16, H=16, geom=good, acc=good, m2d=good, dsh=good, drot=good, start=0, count=-1,
                 rank=0, packed=None)
        a.update(kw)
        return l.gcr_rasterizer_backward_geometry(
            a["P"], a["D"], a["M"], a["means"], a["shs"], a["scales"], 1.0, a["rot"], a["cov"], a["view"],
            a["proj"], a["campos"], a["W"], a["H"], 1.0, 1.0, None, a["geom"], a["acc"], a["m2d"], None, good,
            good, good, good, a["dsh"], good, a["drot"], 0, a["start"], a["count"], a["rank"], 0, 0, a["packed"], None)

    assert geometry(P=0) == 0
    for kw, msg in [(dict(acc=None), "must not be NULL"), (dict(acc=C.c_void_p(0x10008)), "16-byte aligned"),
                    (dict(geom=None), "must not be NULL"), (dict(m2d=None), "must not be NULL"),
                    (dict(dsh=None), "dL_dsh"), (dict(M=25), "at most 16 SH"), (dict(start=3, count=2), "out of bounds"),
                    (dict(scales=None), "scale/rotation"), (dict(drot=C.c_void_p(0x10004)), "dL_drot"),
                    (dict(W=0), "image size"), (dict(rank=16), "shard"),
                    (dict(packed=C.c_void_p(0x10010)), "32-byte aligned")]:
        assert geometry(**kw) < 0, kw
        assert msg in _cabi.last_error(), (kw, _cabi.last_error())

    bw = l.gcr_rasterizer_backward(4, 0, 1, 8, good, 16, 16, good, good, None, good, 1.0, good, None, good, good,
                                   good, 1.0, 1.0, None, good, good, good, good, good, None, good, good, good, good,
                                   good, good, good, 0, 1, 2, None)
    assert bw < 0 and "single-stripe" in _cabi.last_error()
    assert l.gcr_stripe_partition(4, good, good, 1.0, good, None, good, good, 16, 16, 1.0, 1.0, 2, None, good, None) < 0
    assert l.gcr_peer_barrier(None, 0, 2, 1, None) < 0


def test_programmatic_launch_switch_round_trips(built_lib):
    """gcr_set_programmatic_launch returns the previous setting (host-side flag, no device)."""
    from gaussiancity_b200 import _cabi
    first = _cabi.set_programmatic_launch(False)
    assert _cabi.set_programmatic_launch(True) is False
    assert _cabi.set_programmatic_launch(first) is True
    assert _cabi.set_programmatic_launch(first) is first
