"""Script front ends with the reference's command lines (stereo_rerender.py, 3d_view_depthfile.py,
convert_metric_depth_video_to_other_format.py, find_convergence_depth.py); the repo root holds same-named
launchers so `python stereo_rerender.py ...` keeps working."""
