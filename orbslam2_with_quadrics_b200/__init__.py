"""B200-native drop-in for ORB_SLAM2::ORBextractor (yxqc/ORBSLAM2_with_quadrics).

Scope: the per-frame ORB feature front end only (reference src/ORBextractor.cc); see
DESIGN.md.  The product is the CUDA library liborbx.so behind include/orbx.h; this package
holds its sources (csrc/), the C++ adapter that ORB-SLAM2 links (cpp/) and a Python mirror
of the reference class used by tests and benchmarks."""
from .extractor import ORBextractor, Vocabulary  # noqa: F401
from ._capi import OrbxError, KP_DTYPE  # noqa: F401

__all__ = ["ORBextractor", "Vocabulary", "OrbxError", "KP_DTYPE"]
