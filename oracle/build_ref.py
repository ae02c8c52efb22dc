#!/usr/bin/env python
"""Build the UNMODIFIED reference extension into oracle/_ref/ (test infrastructure).

Compiles the five reference source files where they lie under /root/reference
(extensions/diff_gaussian_rasterization/{cuda_rasterizer/{forward,backward,
rasterizer_impl}.cu, rasterize_points.cu, bindings.cpp}) for sm_100a with the
reference's own flags (setup.py:20-37: only `-I third_party/glm`; no
-use_fast_math, default -fmad=true).  The single deviation is `-include cstdint`
(rasterizer_impl.h:24,40-59 uses std::uintptr_t/uint32_t without including
<cstdint>, which gcc 13 rejects).  No reference source is copied into this repo:
objects go to a temp dir, only the linked .so lands in oracle/_ref/ (git-ignored,
but shipped to the GPU box by gpurun).

The result, oracle/_ref/diff_gaussian_rasterization_ext*.so, is the parity pin
for the `-m gpu` tests and the `bench.py --impl reference` arm.  It is never
imported by the product package.
"""
import os
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
REF = os.environ.get("GCR_REFERENCE_ROOT", "/root/reference")
DGR = os.path.join(REF, "extensions", "diff_gaussian_rasterization")
MODNAME = "diff_gaussian_rasterization_ext"


def so_path():
    return os.path.join(OUT_DIR, MODNAME + sysconfig.get_config_var("EXT_SUFFIX"))


def build(force=False, verbose=False):
    """Returns the .so path, or None when /root/reference is absent (GPU box)."""
    out = so_path()
    if not os.path.isdir(DGR):
        return out if os.path.exists(out) else None
    srcs = [
        os.path.join(DGR, "cuda_rasterizer", "rasterizer_impl.cu"),
        os.path.join(DGR, "cuda_rasterizer", "forward.cu"),
        os.path.join(DGR, "cuda_rasterizer", "backward.cu"),
        os.path.join(DGR, "rasterize_points.cu"),
        os.path.join(DGR, "bindings.cpp"),
    ]
    if not force and os.path.exists(out):
        newest = max(os.path.getmtime(s) for s in srcs + [__file__])
        if os.path.getmtime(out) >= newest:
            return out
    import torch  # noqa: F401  (only for include/lib discovery)
    from torch.utils import cpp_extension as ce

    os.makedirs(OUT_DIR, exist_ok=True)
    incs = ce.include_paths() + [sysconfig.get_paths()["include"],
                                 os.path.join(DGR, "third_party", "glm"),
                                 os.path.join(DGR, "cuda_rasterizer")]
    inc_flags = [f"-I{p}" for p in incs]
    common = ["-O3", "-std=c++17", "-include", "cstdint",
              f"-DTORCH_EXTENSION_NAME={MODNAME}", "-DTORCH_API_INCLUDE_EXTENSION_H",
              "-D_GLIBCXX_USE_CXX11_ABI=1"]
    tmp = tempfile.mkdtemp(prefix="gcr_refbuild_")
    objs = []
    procs = []
    for s in srcs:
        o = os.path.join(tmp, os.path.basename(s) + ".o")
        objs.append(o)
        if s.endswith(".cu"):
            cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
                   "-Xcompiler", "-fPIC", "-w"] + common + inc_flags + ["-c", s, "-o", o]
        else:
            cmd = ["g++", "-fPIC", "-w"] + common + inc_flags + ["-c", s, "-o", o]
        if verbose:
            print(" ".join(cmd))
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cmd, p in procs:
        log, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(log.decode())
            raise RuntimeError("reference build failed: " + " ".join(cmd))
    libdirs = ce.library_paths()
    link = ["g++", "-shared", "-o", out] + objs
    for d in libdirs:
        link += [f"-L{d}", f"-Wl,-rpath,{d}"]
    link += ["-L/usr/local/cuda/lib64", "-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python",
             "-lc10_cuda", "-ltorch_cuda", "-lcudart"]
    if verbose:
        print(" ".join(link))
    subprocess.check_call(link)
    write_build_info(common)
    return out


def build_info_path():
    return os.path.join(OUT_DIR, "BUILD_INFO.json")


def write_build_info(flags):
    """The CUDA path reproduces the FMA contraction THIS nvcc chose for the reference sources
    (DESIGN.md section 2); record the toolchain next to the .so so a parity failure on a pin built
    by another compiler can be told from a regression."""
    import json
    try:
        ver = subprocess.run(["nvcc", "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-2:]
    except Exception as e:  # pragma: no cover
        ver = [repr(e)]
    info = {"nvcc": ver, "arch": "sm_100a", "flags": flags, "deviation": "-include cstdint",
            "sources": "extensions/diff_gaussian_rasterization (unmodified)"}
    with open(build_info_path(), "w") as fh:
        json.dump(info, fh, indent=1)


def read_build_info():
    import json
    try:
        with open(build_info_path()) as fh:
            return json.load(fh)
    except Exception:
        return None


# ---- BASELINE config 5 (models/generator.py G-step, ours vs REF): what the harness needs ---------
GE_MODNAME = "grid_encoder_ext"
STAGE_DIR = os.path.join(os.path.dirname(HERE), "baseline", "_ref", "GaussianCity")
STAGED_FILES = [   # the reference's own Python, unmodified, for tools/config5_gstep.py on the GPU box
    "extensions/__init__.py", "extensions/diff_gaussian_rasterization/__init__.py",
    "extensions/grid_encoder/__init__.py", "models/__init__.py", "models/generator.py", "models/pt_v3.py",
    "utils/__init__.py", "utils/helpers.py",
]


def grid_encoder_so_path():
    return os.path.join(OUT_DIR, GE_MODNAME + sysconfig.get_config_var("EXT_SUFFIX"))


def build_grid_encoder(force=False, verbose=False):
    """The reference's other CUDA extension on the generator's import path
    (extensions/grid_encoder/{grid_encoder_ext.cu,bindings.cpp}), compiled as is with the flags of
    its setup.py:29-36 for sm_100a into oracle/_ref/.  Only needed so that the unmodified
    models/generator.py imports; this repo does not rebuild it."""
    out = grid_encoder_so_path()
    ge = os.path.join(REF, "extensions", "grid_encoder")
    if not os.path.isdir(ge):
        return out if os.path.exists(out) else None
    srcs = [os.path.join(ge, "grid_encoder_ext.cu"), os.path.join(ge, "bindings.cpp")]
    if not force and os.path.exists(out) and os.path.getmtime(out) >= max(os.path.getmtime(s) for s in srcs + [__file__]):
        return out
    import torch  # noqa: F401
    from torch.utils import cpp_extension as ce
    os.makedirs(OUT_DIR, exist_ok=True)
    incs = [f"-I{p}" for p in ce.include_paths() + [sysconfig.get_paths()["include"]]]
    common = ["-O3", "-std=c++17", f"-DTORCH_EXTENSION_NAME={GE_MODNAME}", "-DTORCH_API_INCLUDE_EXTENSION_H",
              "-D_GLIBCXX_USE_CXX11_ABI=1"]
    tmp = tempfile.mkdtemp(prefix="gcr_gebuild_")
    objs, procs = [], []
    for s in srcs:
        o = os.path.join(tmp, os.path.basename(s) + ".o")
        objs.append(o)
        if s.endswith(".cu"):
            cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-w",
                   "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
                   "-U__CUDA_NO_HALF2_OPERATORS__"] + common + incs + ["-c", s, "-o", o]
        else:
            cmd = ["g++", "-fPIC", "-w"] + common + incs + ["-c", s, "-o", o]
        if verbose:
            print(" ".join(cmd))
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cmd, p in procs:
        log, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(log.decode())
            raise RuntimeError("grid_encoder build failed: " + " ".join(cmd))
    link = ["g++", "-shared", "-o", out] + objs
    for d in ce.library_paths():
        link += [f"-L{d}", f"-Wl,-rpath,{d}"]
    link += ["-L/usr/local/cuda/lib64", "-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python", "-lc10_cuda",
             "-ltorch_cuda", "-lcudart"]
    subprocess.check_call(link)
    return out


def stage_reference_python():
    """Copy the reference's own Python files the config-5 harness imports into baseline/_ref/
    (git-ignored, shipped to the GPU box, where /root/reference does not exist).  Returns the
    staging root, or None when neither the sources nor a staged copy exist."""
    import shutil
    if not os.path.isdir(REF):
        return STAGE_DIR if os.path.isdir(STAGE_DIR) else None
    for rel in STAGED_FILES:
        src, dst = os.path.join(REF, rel), os.path.join(STAGE_DIR, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
    return STAGE_DIR


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p if p else "reference sources absent and no prebuilt oracle/_ref")
    print(build_grid_encoder(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(stage_reference_python())
