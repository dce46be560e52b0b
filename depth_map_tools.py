"""`import depth_map_tools` as in the reference: the GPU-backed drop-in (see the package module)."""
from metric_depth_video_toolbox_b200.depth_map_tools import *  # noqa: F401,F403
from metric_depth_video_toolbox_b200.depth_map_tools import zero_identity_matrix  # noqa: F401
