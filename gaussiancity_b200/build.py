"""Build the sm_100a CUDA library (libgcr_rasterizer.so) in-tree with nvcc.

No torch headers are involved: the library exposes only the C ABI of
include/gcr_rasterizer.h, so a full rebuild takes seconds.  The .so is git-ignored but travels
to the GPU box with the repo snapshot.
"""
import hashlib
import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
ROOT = os.path.dirname(PKG_DIR)
LIB_PATH = os.path.join(PKG_DIR, "libgcr_rasterizer.so")
STAMP = os.path.join(PKG_DIR, ".libgcr_rasterizer.stamp")

SOURCES = ["api.cu", "preprocess.cu", "binning.cu", "blend_fwd.cu", "blend_bwd.cu",
           "preprocess_bwd.cu", "peer.cu"]
HEADERS = ["gcr_common.cuh", "gcr_kernels.h", "blend_common.cuh",
           os.path.join(ROOT, "include", "gcr_rasterizer.h")]

# Arithmetic flags mirror the reference build (DGR/setup.py:29-36): no -use_fast_math, default
# -fmad=true, IEEE div/sqrt.  -lineinfo so ncu source pages map to these files.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _fingerprint():
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        p = f if os.path.isabs(f) else os.path.join(CSRC, f)
        with open(p, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def build_library(force=False, verbose=False, ptxas_info=False):
    """Compile if sources changed (or force). Returns the .so path."""
    fp = _fingerprint()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == fp:
                return LIB_PATH
    objdir = os.path.join(PKG_DIR, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = nvcc_path()
    procs = []
    objs = []
    for s in SOURCES:
        o = os.path.join(objdir, s + ".o")
        objs.append(o)
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if ptxas_info else []) + \
              ["-c", os.path.join(CSRC, s), "-o", o]
        if verbose:
            print(" ".join(cmd))
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    logs = []
    for cmd, p in procs:
        out, _ = p.communicate()
        logs.append(out.decode())
        if p.returncode != 0:
            sys.stderr.write(out.decode())
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    link = [nvcc, "-shared", "-o", LIB_PATH] + objs + \
           ["-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"]
    if verbose:
        print(" ".join(link))
    subprocess.check_call(link)
    with open(STAMP, "w") as fh:
        fh.write(fp)
    if ptxas_info or verbose:
        print("".join(logs))
    return LIB_PATH


# ---- the hash-grid encoder: its own library behind include/gcr_grid_encoder.h -----------
GRID_LIB_PATH = os.path.join(PKG_DIR, "libgcr_grid_encoder.so")


def build_grid_library(force=False, verbose=False, ptxas_info=False):
    """nvcc, one translation unit (csrc/grid_encoder.cu) -> libgcr_grid_encoder.so, in-tree."""
    src = os.path.join(CSRC, "grid_encoder.cu")
    hdr = os.path.join(ROOT, "include", "gcr_grid_encoder.h")
    stamp = os.path.join(PKG_DIR, ".libgcr_grid_encoder.stamp")
    h = hashlib.sha256()
    for f in (src, hdr):
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    fp = h.hexdigest()
    if not force and os.path.exists(GRID_LIB_PATH) and os.path.exists(stamp) and \
            open(stamp).read().strip() == fp:
        return GRID_LIB_PATH
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if ptxas_info else []) + \
          ["-shared", src, "-o", GRID_LIB_PATH]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    with open(stamp, "w") as fh:
        fh.write(fp)
    return GRID_LIB_PATH


# ---- Seam A: the native module `diff_gaussian_rasterization_ext` (csrc/torch_module.cpp) -----------
MODNAME = "diff_gaussian_rasterization_ext"
COMPAT_DIR = os.path.join(PKG_DIR, "compat")


def native_module_path(modname=MODNAME):
    import sysconfig
    return os.path.join(COMPAT_DIR, modname + sysconfig.get_config_var("EXT_SUFFIX"))


def _build_torch_module(modname, src_name, hdr_name, link_lib, stamp_name, force=False, verbose=False):
    """g++ only (host code over a C ABI; ~1-2 min of torch headers).  The module links its CUDA
    library through an $ORIGIN-relative rpath, so the pair travels together."""
    import sysconfig
    out = native_module_path(modname)
    src = os.path.join(CSRC, src_name)
    hdr = os.path.join(ROOT, "include", hdr_name)
    stamp = os.path.join(COMPAT_DIR, stamp_name)
    h = hashlib.sha256()
    for f in (src, hdr):
        with open(f, "rb") as fh:
            h.update(fh.read())
    import torch
    h.update(torch.__version__.encode())
    fp = h.hexdigest()
    if not force and os.path.exists(out) and os.path.exists(stamp) and open(stamp).read().strip() == fp:
        return out
    from torch.utils import cpp_extension as ce
    os.makedirs(COMPAT_DIR, exist_ok=True)
    incs = ce.include_paths() + [sysconfig.get_paths()["include"], "/usr/local/cuda/include"]
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-w", f"-DTORCH_EXTENSION_NAME={modname}",
           "-DTORCH_API_INCLUDE_EXTENSION_H", "-D_GLIBCXX_USE_CXX11_ABI=1"] + [f"-I{p}" for p in incs] + \
          [src, "-o", out, f"-L{PKG_DIR}", f"-l{link_lib}", "-Wl,-rpath,$ORIGIN/.."]
    for d in ce.library_paths():
        cmd += [f"-L{d}", f"-Wl,-rpath,{d}"]
    cmd += ["-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    with open(stamp, "w") as fh:
        fh.write(fp)
    return out


def build_native_module(force=False, verbose=False):
    """Seam A of the rasterizer: `diff_gaussian_rasterization_ext` over libgcr_rasterizer.so."""
    build_library()
    return _build_torch_module(MODNAME, "torch_module.cpp", "gcr_rasterizer.h", "gcr_rasterizer",
                               ".native.stamp", force, verbose)


GRID_MODNAME = "grid_encoder_ext"


def build_grid_native_module(force=False, verbose=False):
    """Seam A of the grid encoder: `grid_encoder_ext` (csrc/grid_module.cpp) over
    libgcr_grid_encoder.so -- what the reference's extensions/grid_encoder/__init__.py imports."""
    build_grid_library()
    return _build_torch_module(GRID_MODNAME, "grid_module.cpp", "gcr_grid_encoder.h", "gcr_grid_encoder",
                               ".grid_native.stamp", force, verbose)


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv,
                        ptxas_info="--ptxas" in sys.argv))
    print(build_grid_library(force="--force" in sys.argv, verbose="-v" in sys.argv,
                             ptxas_info="--ptxas" in sys.argv))
    if "--native" in sys.argv:
        print(build_native_module(force="--force" in sys.argv, verbose="-v" in sys.argv))
        print(build_grid_native_module(force="--force" in sys.argv, verbose="-v" in sys.argv))
