"""Launcher with the reference script's name and command line (see metric_depth_video_toolbox_b200/cli/convert_format.py)."""
from metric_depth_video_toolbox_b200.cli.convert_format import main

if __name__ == "__main__":
    raise SystemExit(main())
