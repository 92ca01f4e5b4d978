"""svi_mapper_b200 -- B200-native (sm_100a) stereo front-end hot path of svi_mapper:
Harris/GFTT detection, BRIEF-32 description, dense scan-line Hamming matching, triangulation and
projection-window landmark tracking, behind a C-ABI (include/svi_gpu.h -> libsvi_gpu.so)."""
from .calib import PinholeCamera, StereoCamera, construct_camera_stereo, load_camera  # noqa: F401
from .frontend import NoMatchFound, StereoFrames, StereoFrontend, SviError, default_params, status_text  # noqa: F401
from .partition import MultiFrontend, frame_range  # noqa: F401
