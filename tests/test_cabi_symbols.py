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
