"""havatar_b200: B200-native (sm_100a) implementation of the HAvatar volumetric-render hot path.

The CUDA library (havatar_b200/csrc -> libhavatar_b200.so, C ABI declared in include/havatar_b200.h)
is loaded lazily by havatar_b200._lib; there is no CPU fallback -- every op raises if it is missing.
"""
__version__ = "0.1.0"
