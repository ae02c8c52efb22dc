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


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv,
                        ptxas_info="--ptxas" in sys.argv))
