"""mpt_b200 -- B200-native batched nearest-neighbour search and state/edge validity checking behind
the Motion Planning Templates interfaces.  The product is libmptg.so (CUDA, sm_100a) and the C++
host layer under include/mptg/; this package is the thin ctypes mirror used by tests and bench.py.
"""
from . import _lib
from ._lib import (F32, F64, KNN_AUTO, KNN_BRUTE, KNN_BVH, NO_INDEX, MptgError)
from .api import (Comm, Context, DevicePPRM, DevicePRRT, DevicePRRTStar, Nearest, Scenario, Space, knn_merge_dev, lp_space, sample, sample_from_uniforms,
                  se2_space, se3_space, so2_space, so3_space)

__all__ = [
    "Comm", "Context", "DevicePPRM", "DevicePRRT", "DevicePRRTStar", "Nearest", "Scenario", "Space", "knn_merge_dev", "sample", "sample_from_uniforms", "lp_space", "se2_space", "se3_space", "so2_space",
    "so3_space", "F32", "F64", "KNN_AUTO", "KNN_BRUTE", "KNN_BVH", "NO_INDEX", "MptgError",
]
