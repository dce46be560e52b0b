"""`movie_2_3D` with the reference's module surface (MDVT_gui.py:1290-1320 loads the script by path and calls the steps by
name): steps 4 and 5 -- the two that run the dense per-frame path -- are GPU-backed (movie_steps.py); the scene planning
helpers and step 1 are host code with the reference's results (movie_plan.py, pinned by tests/golden/movie_2_3D.json); the
steps that wrap third-party models or ffmpeg (2, 3, 6, 7) are importable and refuse with the reason."""
from metric_depth_video_toolbox_b200.movie_plan import (  # noqa: F401
    _seconds_to_timecode, ensure_output_dir, ensure_scene_file, is_valid_video, load_and_split_scenes, main, open_input_video, parse_args,
    plan_scene_files, split_scenes, step1_create_scene_videos, step2_estimate_depth, step3_generate_masks, step6_infill_and_collect,
    step6_inspatio_world_infill_and_collect, step6_m2svid_infill_and_collect, step6_normal_infill_render_sbs,
    step6_stereo_dissoclusion_net_infill_and_collect, step6_stereocrafter_infill_and_collect, step7_concat_and_mux, validate_video_lengths,
    wait_for_first, write_frames_to_file)
from metric_depth_video_toolbox_b200.movie_steps import step4_find_convergence, step5_render_sbs, stereo_rerender_argv  # noqa: F401

if __name__ == "__main__":
    main()
