"""slam_cin0051_b200 -- B200-native (sm_100a) implementation of the monocular SLAM frontend hot path of
daviyan5/SLAM-CIN0051, behind the reference's src/frontend + src/preprocessing interfaces.

Layout: csrc/ (CUDA kernels + C ABI, built into libslamcu.so), _lib.py (ctypes binding),
frontend.py / common.py / preprocessing.py / pose.py (host-side mirrors of the reference classes),
sequence.py (device-resident batched path).  No CPU fallback anywhere.
"""
from ._lib import KEYPOINT_DTYPE, KNN2_DTYPE, MATCH_DTYPE, Context, SlamcuError, load  # noqa: F401
from .common import Camera, bgr_to_gray  # noqa: F401
from .frontend import FeatureDetector, FeatureMatcher  # noqa: F401
from .pose import PoseEstimator, estimate_pose, find_essential, fivept_solve, ransac_score  # noqa: F401
from .sequence import FrameSequence  # noqa: F401

__all__ = ["Context", "SlamcuError", "FeatureDetector", "FeatureMatcher", "FrameSequence", "Camera", "bgr_to_gray", "PoseEstimator", "find_essential",
           "KEYPOINT_DTYPE", "MATCH_DTYPE", "KNN2_DTYPE", "load"]
