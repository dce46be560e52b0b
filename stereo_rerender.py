"""Launcher with the reference script's name and command line; also re-exports the helpers other reference
modules import from `stereo_rerender` (basic_nomal_infill.py:10)."""
from metric_depth_video_toolbox_b200.cli.stereo_rerender import *  # noqa: F401,F403
from metric_depth_video_toolbox_b200.cli.stereo_rerender import main

if __name__ == "__main__":
    raise SystemExit(main())
