"""rgbid_slam_b200 -- B200-native dense frame-to-keyframe alignment (the hot path of dangut/RGBiD-SLAM).

Layout: csrc/ (hand-written sm_100a CUDA kernels + the C ABI of include/rgbid_b200.h), capi.py (ctypes
binding of that ABI), host.py (torch-tensor front end used by tests and bench.py), synth.py (synthetic
TUM-format RGB-D scenes).  The directory is named `rgbid-slam_b200`; import it as `rgbid_slam_b200`
(the repository root holds a one-file loader for that name).
"""
from . import capi  # noqa: F401
from .build import build, LIB  # noqa: F401


def load():
    """Load the CUDA library (raises if it has not been built: there is no CPU fallback)."""
    return capi.load()
