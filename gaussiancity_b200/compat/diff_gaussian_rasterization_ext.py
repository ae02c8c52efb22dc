"""Drop-in module with the reference's native-extension name (DGR/__init__.py:16 imports
`diff_gaussian_rasterization_ext`).  Put this directory on sys.path and the reference's
UNMODIFIED extensions/diff_gaussian_rasterization/__init__.py runs on the B200 kernels."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from gaussiancity_b200.ext import (mark_visible, rasterize_gaussians,  # noqa: E402,F401
                                   rasterize_gaussians_backward)
