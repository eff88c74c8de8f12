"""Drop-in for the reference's `model/op` package (reference model/op/__init__.py:1-2): same public
names.  `install_reference_modules()` additionally registers the two bare-name extension modules the
UNMODIFIED reference wrappers import (`import fused`, `import upfirdn2d`: model/op/fused_act.py:20,
model/op/upfirdn2d.py:19), so the reference tree runs on these kernels without being edited."""
import sys

from . import conv2d_gradfix, fused, upfirdn2d_op
from .fused_act import FusedLeakyReLU, fused_leaky_relu
from .upfirdn2d import upfirdn2d


def install_reference_modules():
    sys.modules["fused"] = fused
    sys.modules["upfirdn2d"] = upfirdn2d_op


__all__ = ["FusedLeakyReLU", "fused_leaky_relu", "upfirdn2d", "conv2d_gradfix", "install_reference_modules"]
