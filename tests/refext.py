"""Test-only helpers around the UNMODIFIED reference extension built into oracle/_ref/ by
oracle/build_ref.py: loading it, calling it on a Scene, and reading its opaque buffers
(layout = GeometryState/ImageState/BinningState::fromChunk, DGR rasterizer_impl.cu:134-174,
128-byte aligned sections; only the arrays in front of the cub temp storage are located)."""
import glob
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def load_reference_ext():
    """Returns the reference's pybind module or None if oracle/_ref has no build."""
    if not glob.glob(os.path.join(REF_DIR, "diff_gaussian_rasterization_ext*.so")):
        return None
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import importlib
    return importlib.import_module("diff_gaussian_rasterization_ext")


def _al(x, a=128):
    return (x + a - 1) // a * a


def ref_geom_views(geom, P):
    off = 0
    out = {}

    def take(name, nbytes, dtype, shape):
        nonlocal off
        off = _al(off)
        out[name] = geom[off:off + nbytes].view(dtype).view(*shape)
        off += nbytes

    take("depths", 4 * P, torch.float32, (P,))
    take("clamped", 3 * P, torch.bool, (P, 3))
    take("internal_radii", 4 * P, torch.int32, (P,))
    take("means2D", 8 * P, torch.float32, (P, 2))
    take("cov3D", 24 * P, torch.float32, (P, 6))
    take("conic_opacity", 16 * P, torch.float32, (P, 4))
    take("rgb", 12 * P, torch.float32, (P, 3))
    take("tiles_touched", 4 * P, torch.int32, (P,))
    return out


def ref_binning_views(binning, R):
    off = 0
    out = {}

    def take(name, nbytes, dtype, shape):
        nonlocal off
        off = _al(off)
        out[name] = binning[off:off + nbytes].view(dtype).view(*shape)
        off += nbytes

    take("point_list", 4 * R, torch.int32, (R,))
    take("point_list_unsorted", 4 * R, torch.int32, (R,))
    take("point_list_keys", 8 * R, torch.int64, (R,))
    take("point_list_keys_unsorted", 8 * R, torch.int64, (R,))
    return out


def ref_img_views(img, H, W):
    N = H * W
    off = 0
    out = {}

    def take(name, nbytes, dtype, shape):
        nonlocal off
        off = _al(off)
        out[name] = img[off:off + nbytes].view(dtype).view(*shape)
        off += nbytes

    take("accum_alpha", 4 * N, torch.float32, (H, W))
    take("n_contrib", 4 * N, torch.int32, (H, W))
    take("ranges", 8 * N, torch.int32, (N, 2))
    return out


def our_views(P, R, W, H, geom, binning, img):
    """Views into OUR opaque buffers through gcr_debug_offset (include/gcr_rasterizer.h)."""
    from gaussiancity_b200 import _cabi as C
    lib = C.lib()
    off = lambda which: lib.gcr_debug_offset(which, P, R, W, H)
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    v = {}
    v["records"] = geom[off(C.GEOM_RECORDS):off(C.GEOM_RECORDS) + 48 * P].view(torch.float32).view(P, 12)
    v["tiles_touched"] = geom[off(C.GEOM_TILES_TOUCHED):off(C.GEOM_TILES_TOUCHED) + 4 * P].view(torch.int32)
    v["clamped"] = geom[off(C.GEOM_CLAMPED):off(C.GEOM_CLAMPED) + P]
    v["sorted_gauss"] = geom[off(C.GEOM_SORTED_GAUSS):off(C.GEOM_SORTED_GAUSS) + 4 * P].view(torch.int32)
    v["sorted_depth_keys"] = geom[off(C.GEOM_DEPTH_SORTED_KEYS):off(C.GEOM_DEPTH_SORTED_KEYS) + 4 * P].view(torch.int32)
    v["offsets"] = geom[off(C.GEOM_OFFSETS):off(C.GEOM_OFFSETS) + 4 * P].view(torch.int32)
    v["owner"] = geom[off(C.GEOM_OWNER):off(C.GEOM_OWNER) + P]
    # device-side counters: [0] Gaussians that reach this rank's stripe (= entries of sorted_gauss /
    # sorted_depth_keys / offsets that are valid), [1] num_rendered, [2] overflow flag
    v["counters"] = geom[off(C.GEOM_COUNTERS):off(C.GEOM_COUNTERS) + 12].view(torch.int32)
    if R > 0:
        v["point_list"] = binning[off(C.BIN_POINT_LIST):off(C.BIN_POINT_LIST) + 4 * R].view(torch.int32)
        v["tile_keys"] = binning[off(C.BIN_TILE_KEYS):off(C.BIN_TILE_KEYS) + 4 * R].view(torch.int32)
    v["final_T"] = img[off(C.IMG_FINAL_T):off(C.IMG_FINAL_T) + 4 * W * H].view(torch.float32).view(H, W)
    v["n_contrib"] = img[off(C.IMG_N_CONTRIB):off(C.IMG_N_CONTRIB) + 4 * W * H].view(torch.int32).view(H, W)
    v["ranges"] = img[off(C.IMG_RANGES):off(C.IMG_RANGES) + 8 * tiles].view(torch.int32).view(tiles, 2)
    return v


def scene_forward_args(s, scale_modifier=1.0):
    """Positional args of rasterize_gaussians (DGR/bindings.cpp:16) for a Scene."""
    e = torch.Tensor([])
    return (s.bg, s.means3D, s.colors_precomp if s.colors_precomp is not None else e, s.opacities,
            s.scales, s.rotations, scale_modifier, e, s.view_matrix, s.proj_matrix, s.tanfovx, s.tanfovy,
            s.img_h, s.img_w, s.shs if s.shs is not None else e, s.sh_degree, s.campos, False, False)


def scene_backward_args(s, radii, grad_out, geom, R, binning, img, scale_modifier=1.0):
    e = torch.Tensor([])
    return (s.bg, s.means3D, radii, s.colors_precomp if s.colors_precomp is not None else e,
            s.scales, s.rotations, scale_modifier, e, s.view_matrix, s.proj_matrix, s.tanfovx, s.tanfovy,
            grad_out, s.shs if s.shs is not None else e, s.sh_degree, s.campos, geom, R, binning,
            img, False)


# ---- grid encoder ---------------------------------------------------------------------
def _load_extension_file(name, path):
    """Load a pybind .so by path without registering it in sys.modules (two modules called
    `grid_encoder_ext` -- the reference's and ours -- live in one process)."""
    import importlib.machinery
    import importlib.util
    loader = importlib.machinery.ExtensionFileLoader(name, path)
    spec = importlib.util.spec_from_file_location(name, path, loader=loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    return mod


def load_reference_grid_ext():
    """The UNMODIFIED reference grid_encoder_ext built by oracle/build_ref.build_grid_encoder(), or None."""
    hits = glob.glob(os.path.join(REF_DIR, "grid_encoder_ext*.so"))
    return _load_extension_file("grid_encoder_ext", hits[0]) if hits else None


def load_our_grid_ext():
    from gaussiancity_b200 import build
    return _load_extension_file("grid_encoder_ext", build.native_module_path("grid_encoder_ext"))


def load_reference_grid_python(ext_module, alias):
    """The reference's own extensions/grid_encoder/__init__.py (staged under baseline/_ref by
    oracle/build_ref.py, or /root/reference where it exists), executed with `import grid_encoder_ext`
    resolving to `ext_module`.  Returns the module (GridEncoder, GridEncoderFunction) or None."""
    import importlib.util
    for root in (os.path.join(ROOT, "baseline", "_ref", "GaussianCity"), "/root/reference"):
        path = os.path.join(root, "extensions", "grid_encoder", "__init__.py")
        if os.path.exists(path):
            break
    else:
        return None
    saved = sys.modules.get("grid_encoder_ext")
    sys.modules["grid_encoder_ext"] = ext_module
    try:
        spec = importlib.util.spec_from_file_location(alias, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if saved is None:
            sys.modules.pop("grid_encoder_ext", None)
        else:
            sys.modules["grid_encoder_ext"] = saved
    return mod
