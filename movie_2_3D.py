"""`from movie_2_3D import step4_find_convergence, step5_render_sbs` as in the reference (MDVT_gui.py:1290-1320 imports
the steps by name): the two steps of the pipeline that run the dense per-frame path, GPU-backed.  The other steps
(scene split, depth models, masks, learned infill, muxing) are out of scope here -- see the package module."""
from metric_depth_video_toolbox_b200.movie_steps import is_valid_video, step4_find_convergence, step5_render_sbs, stereo_rerender_argv  # noqa: F401
