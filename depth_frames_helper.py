"""`import depth_frames_helper` as in the reference: the GPU-backed drop-in (see the package module)."""
from metric_depth_video_toolbox_b200.depth_frames_helper import *  # noqa: F401,F403
from metric_depth_video_toolbox_b200.depth_frames_helper import A, C  # noqa: F401
