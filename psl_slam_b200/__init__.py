"""psl_slam_b200 — B200-native feature front end of PSL-SLAM (ORB + lines + matching).

Host-side mirrors of the reference's extractor / matcher classes over the C-ABI in
include/psl_frontend.h (libpsl_frontend.so, hand-written sm_100a CUDA).  No CPU fallback.
"""
from ._lib import KEYLINE_DTYPE, KP_DTYPE, PslError, default_config  # noqa: F401
from .orb import Context, ORBextractor  # noqa: F401
from .line import LINEextractor  # noqa: F401
from .line_matcher import InsectLineMatch, LineFrameData, LSDmatcher, line_junctions, lines_3d, plane_hypotheses  # noqa: F401
from .matcher import FrameData, ORBmatcher, hamming_knn2  # noqa: F401
from .tracking import (convert_rgbd, image_bounds, make_camera, make_distortion, make_track_params,  # noqa: F401
                       track_frontend_batch, track_frontend_batch_dev, track_orb_batch, track_orb_batch_dev,
                       undistort_keypoints)
from .optimizer import PoseOptimization  # noqa: F401,E402
