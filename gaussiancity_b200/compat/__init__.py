"""Seam A: put THIS directory on sys.path and the reference's unmodified
extensions/diff_gaussian_rasterization/__init__.py imports `diff_gaussian_rasterization_ext`
(DGR/__init__.py:16) from here -- the native torch module built by
gaussiancity_b200/build.py::build_native_module() over the C ABI of include/gcr_rasterizer.h.
The package itself imports the same module as gaussiancity_b200.compat.diff_gaussian_rasterization_ext."""
