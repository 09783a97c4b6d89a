"""linemod_pose_estimation_b200 -- B200-native LINEMOD matcher behind the cv::linemod::Detector surface.

The product is the C-ABI library liblinemod_b200.so (csrc/, include/linemod_b200.h): hand-written sm_100a CUDA
kernels for the path cv::linemod::Detector::match, which the reference enters at
/root/reference/src/rgbdDetector.cpp:31-34.  `Detector` is the Python mirror of that class.
"""
from .detector import (ColorGradient, DepthNormal, Detector, DetectorGroup, FrameStream, QuantizedPyramid, Stage,  # noqa: F401
                       process)
from ._capi import LinemodError, MATCH_DTYPE, RAW_DTYPE  # noqa: F401
from .training import Mesh, ViewSphere, camera  # noqa: F401
