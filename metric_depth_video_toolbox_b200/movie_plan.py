"""Host-side planning helpers of movie_2_3D.py with the reference's names and results (MDVT_gui.py loads the script by path
and calls them; `tests/golden/movie_2_3D.json` holds the reference's own outputs): scene splitting, per-scene file plan,
time codes, the PySceneDetect csv loader, the flag table, scene clip writing (step 1), and -- importable, refusing with the
reason -- the steps that wrap third-party models or ffmpeg (2, 3, 6, 7), which are outside the dense per-frame path
(SURVEY.md 8).  Nothing here touches a pixel of the hot path; steps 4 and 5 live in movie_steps.py."""
from __future__ import annotations

import argparse
import csv
import os
import shutil
import subprocess
import time
from typing import Dict, List


def _seconds_to_timecode(seconds: float) -> str:
    """HH:MM:SS.mmm of a time in seconds, rounded to the millisecond (movie_2_3D.py:103-108)."""
    total_ms = round(seconds * 1000)
    hours, rest = divmod(total_ms, 3_600_000)
    minutes, rest = divmod(rest, 60_000)
    secs, ms = divmod(rest, 1000)
    return f"{hours:02d}:{minutes:02d}:{secs:02d}.{ms:03d}"


def _piece(scene: Dict, first: int, last: int, start_frame: int, start_s: float, seconds_per_frame: float) -> Dict:
    """One output row of split_scenes: the input row's extra keys kept, the frame / time fields recomputed for [first, last]."""
    begin = start_s + (first - start_frame) * seconds_per_frame
    end = start_s + (last - start_frame) * seconds_per_frame
    span = max(0.0, end - begin)
    row = dict(scene)
    row.update({"Scene Number": None,
                "Start Frame": str(first), "Start Time (seconds)": f"{begin:.3f}", "Start Timecode": _seconds_to_timecode(begin),
                "End Frame": str(last), "End Time (seconds)": f"{end:.3f}", "End Timecode": _seconds_to_timecode(end),
                "Length (frames)": str(last - first + 1), "Length (seconds)": f"{span:.3f}", "Length (timecode)": _seconds_to_timecode(span)})
    return row


def split_scenes(scenes, max_scene_frames: int = 1500):
    """Scenes longer than `max_scene_frames` cut into consecutive pieces of at most that length, every row's fields
    normalised, 'Scene Number' renumbered from 1 (movie_2_3D.py:111-173).  Returns a new list of dicts."""
    rows = []
    for scene in scenes:
        first, last = int(scene["Start Frame"]), int(scene["End Frame"])
        start_s, end_s = float(scene["Start Time (seconds)"]), float(scene["End Time (seconds)"])
        spf = (end_s - start_s) / (last - first) if last != first else 0.0   # constant frame rate
        n = last - first + 1
        if n <= max_scene_frames:   # also the degenerate n <= 0 rows: kept, normalised
            rows.append(_piece(scene, first, last, first, start_s, spf))
            continue
        for a in range(first, last + 1, max_scene_frames):
            rows.append(_piece(scene, a, min(a + max_scene_frames - 1, last), first, start_s, spf))
    for number, row in enumerate(rows, start=1):
        row["Scene Number"] = str(number)
    return rows


def parse_args() -> argparse.Namespace:
    """The command line of movie_2_3D.py (:180-201): same flags, types, defaults and the `need --color_video` rule."""
    p = argparse.ArgumentParser(description="Takes a movie and converts it in to stereo 3D")
    p.add_argument("--color_video", type=str, default=None, required=False, help="video file to use as color input")
    p.add_argument("--scene_file", type=str, default=None, required=False, help="csv from PySceneDetect describing the scenes")
    p.add_argument("--csv_delimiter", type=str, default=",", required=False, help="Delimiter used in csv")
    p.add_argument("--output_dir", type=str, default="output", required=False, help="folder where output will be placed")
    p.add_argument("--end_scene", type=int, default=-1, required=False, help="Stop after a certain scene nr")
    p.add_argument("--no_render", action="store_true", required=False, help="Skip rendering and subseqvent steps.")
    p.add_argument("--parallel", type=int, default=int(os.cpu_count() // 2), help="Run some steps in parallel, for faster processing.")
    p.add_argument("--max_scene_frames", type=int, default=1500, help="Max length of scene in nr of frames, longer scenes will be processed in chunks.")
    p.add_argument("--infill_engine", type=str, default="stereocrafter", required=False,
                   help="What infill engine to use. (none, normals, stereo_dissoclusion_net, stereocrafter, m2svid)")
    p.add_argument("--gui", action="store_true", help="Launch the PySide6 GUI")
    args = p.parse_args()
    if args.color_video is None and not args.gui:
        raise ValueError("need --color_video")
    return args


def ensure_output_dir(path: str) -> None:
    """movie_2_3D.py:204-206."""
    if not os.path.exists(path):
        os.makedirs(path)


def ensure_scene_file(args: argparse.Namespace) -> None:
    """Without --scene_file: `<output_dir>/<video name>-Scenes.csv`, produced by PySceneDetect's command line tool when it
    does not exist yet (movie_2_3D.py:209-222)."""
    if args.scene_file is not None:
        return
    name = os.path.splitext(os.path.basename(args.color_video))[0] + "-Scenes.csv"
    args.scene_file = os.path.join(args.output_dir, name)
    if not os.path.exists(args.scene_file):
        subprocess.run(f"scenedetect -i {args.color_video} list-scenes", shell=True)
        shutil.move(name, args.scene_file)


def open_input_video(color_video_path: str):
    """-> (cv2.VideoCapture, width, height, frame rate) (movie_2_3D.py:225-230)."""
    import cv2

    cap = cv2.VideoCapture(color_video_path)
    return cap, int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)), int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT)), cap.get(cv2.CAP_PROP_FPS)


def load_and_split_scenes(scene_csv_path: str, csv_delimiter: str, max_scene_frames: int) -> List[Dict]:
    """PySceneDetect's scene list (first line: a time code list, skipped; then a header row and one row per scene) through
    split_scenes (movie_2_3D.py:233-241)."""
    with open(scene_csv_path, newline="") as fh:
        next(csv.reader(fh))
        rows = list(csv.DictReader(fh, delimiter=csv_delimiter))
    return split_scenes(rows, max_scene_frames=max_scene_frames)


def plan_scene_files(scenes: List[Dict], output_dir: str, end_scene: int) -> List[Dict]:
    """Per-scene file names and switches, in place and returned, up to and including scene number `end_scene`
    (movie_2_3D.py:244-280): scene_<n>.mkv and its _depth / _mask / _xfovs / _stereo / _infillmask / _infilled companions;
    'infill' / 'convergence' are on unless the row says 'No'; 'finished' when the stereo or the infilled video exists."""
    planned = []
    for scene in scenes:
        base = os.path.join(output_dir, f"scene_{scene['Scene Number']}.mkv")
        depth = base + "_depth.mkv"
        sbs = depth + "_stereo.mkv"
        scene.update(scene_video_file=base, depth_video_file=depth, mask_video_file=base + "_mask.mkv", xfovs_file=depth + "_xfovs.json",
                     sbs=sbs, sbs_infill=sbs + "_infillmask.mkv", infilled=sbs + "_infilled.mkv")
        scene["infill"] = scene.get("Infill") != "No"
        scene["convergence"] = scene.get("Convergence") != "No"
        scene["finished"] = os.path.exists(scene["sbs"]) or os.path.exists(scene["infilled"])
        planned.append(scene)
        if end_scene == int(scene["Scene Number"]):
            break
    return planned


def is_valid_video(file_path) -> bool:
    """The file exists and holds at least 2 KB (movie_2_3D.py:62-67)."""
    return os.path.exists(file_path) and os.path.getsize(file_path) >= 2048


def validate_video_lengths(scene_video_files) -> bool:
    """Every scene's 'infilled' video exists, opens and holds 'Length (frames)' frames (movie_2_3D.py:70-100)."""
    import cv2

    problems = []
    for scene in scene_video_files:
        path, expected = scene["infilled"], int(scene["Length (frames)"])
        if not os.path.isfile(path):
            print(f"File does not exist: {path}")
            problems.append((path, "File not found"))
            continue
        cap = cv2.VideoCapture(path)
        if not cap.isOpened():
            print(f"Could not open file: {path}")
            problems.append((path, "Could not open"))
            continue
        actual = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
        cap.release()
        if actual != expected:
            print(f"Mismatch in {path}: expected {expected}, got {actual}")
            problems.append((path, f"Expected {expected}, got {actual}"))
    if problems:
        print(f"Some files had issues delete them and run again: {problems}")
    return not problems


def wait_for_first(processes):
    """Blocks until one of the subprocess.Popen objects has finished; returns the list without it (movie_2_3D.py:41-59)."""
    if not processes:
        return []
    while True:
        for k, proc in enumerate(processes):
            if proc.poll() is not None:
                return processes[:k] + processes[k + 1:]
        time.sleep(0.1)


def write_frames_to_file(input_video, nr_frames_to_copy, scene_video_file, frame_rate, frame_width, frame_height):
    """Reads `nr_frames_to_copy` frames from the open capture (at least one read, like the reference's loop) and, when a file
    name is given, writes them as FFV1 (movie_2_3D.py:21-38); with scene_video_file None the frames are skipped."""
    import cv2

    out = None
    if scene_video_file is not None:
        out = cv2.VideoWriter(scene_video_file, cv2.VideoWriter_fourcc(*"FFV1"), frame_rate, (frame_width, frame_height))
    copied = 0
    while input_video.isOpened():
        ok, frame = input_video.read()
        if not ok:
            break
        if out is not None:
            out.write(frame)
        copied += 1
        if copied >= nr_frames_to_copy:
            break
    if out is not None:
        out.release()


def step1_create_scene_videos(raw_video, scene_video_files: List[Dict], frame_rate, frame_width, frame_height) -> None:
    """Per-scene FFV1 clips copied frame by frame from the input video; scenes that exist or are finished are read past
    (movie_2_3D.py:283-301)."""
    from . import depth_frames_helper

    print("Step one: create video files for all scenes")
    if not any(not s["finished"] and not os.path.exists(s["scene_video_file"]) for s in scene_video_files):
        return
    for scene in scene_video_files:
        print("scene:", str(scene["Scene Number"]))
        tmp = None
        if not scene["finished"] and not os.path.exists(scene["scene_video_file"]):
            tmp = str(scene["scene_video_file"]) + "_tmp.mkv"
            print("create:", str(scene["scene_video_file"]))
        write_frames_to_file(raw_video, int(scene["Length (frames)"]), tmp, frame_rate, frame_width, frame_height)
        if tmp is not None:
            depth_frames_helper.verify_and_move(tmp, int(scene["Length (frames)"]), scene["scene_video_file"])


class OutOfScope(NotImplementedError):
    """A step of movie_2_3D.py that wraps third-party models or tools outside the dense per-frame path."""


def _refuse(step: str, what: str):
    raise OutOfScope(f"movie_2_3D.{step}: {what} -- outside the dense per-frame path this package implements (SURVEY.md section 8); "
                     f"run that step with the reference's own script, then steps 4 and 5 here")


def step2_estimate_depth(args, scene_video_files) -> None:
    """movie_2_3D.py:304-384 drives the depth-model wrappers (video_metric_convert.py, unik3d / geometrycrafter ...)."""
    _refuse("step2_estimate_depth", "depth estimation runs third-party model wrappers")


def step3_generate_masks(args, scene_video_files) -> None:
    """movie_2_3D.py:387-405 drives the focus-mask model (generate_video_mask.py)."""
    _refuse("step3_generate_masks", "focus masks come from a third-party segmentation model")


def step6_normal_infill_render_sbs(args, scene_video_files):
    _refuse("step6_normal_infill_render_sbs", "drives basic_nomal_infill.py per scene (its kernel is here: stereo_rerender.infill_using_normals / --do_basic_infill)")


def step6_m2svid_infill_and_collect(args, scene_video_files):
    _refuse("step6_m2svid_infill_and_collect", "learned infill (m2svid)")


def step6_inspatio_world_infill_and_collect(args, scene_video_files):
    _refuse("step6_inspatio_world_infill_and_collect", "learned infill (inspatio world)")


def step6_stereocrafter_infill_and_collect(args, scene_video_files):
    _refuse("step6_stereocrafter_infill_and_collect", "learned infill (StereoCrafter)")


def step6_stereo_dissoclusion_net_infill_and_collect(args, scene_video_files):
    _refuse("step6_stereo_dissoclusion_net_infill_and_collect", "learned infill (stereo dissoclusion net)")


def step6_infill_and_collect(args, scene_video_files):
    """The dispatcher of movie_2_3D.py:454-469: same engine names, same error for an unknown one."""
    engines = {"normals": step6_normal_infill_render_sbs, "stereo_dissoclusion_net": step6_stereo_dissoclusion_net_infill_and_collect,
               "stereocrafter": step6_stereocrafter_infill_and_collect, "m2svid": step6_m2svid_infill_and_collect,
               "inspatio_world": step6_inspatio_world_infill_and_collect}
    if args.infill_engine not in engines:
        raise Exception(f"unknown infill engine: {args.infill_engine}")
    return engines[args.infill_engine](args, scene_video_files)


def step7_concat_and_mux(args, video_files_to_concat) -> None:
    """movie_2_3D.py:702-782 concatenates the scenes and muxes the audio with ffmpeg."""
    _refuse("step7_concat_and_mux", "concatenation and audio muxing are ffmpeg command lines")


def main():
    """movie_2_3D.py:785-: the whole pipeline needs steps 2, 3, 6 and 7."""
    _refuse("main", "the full pipeline includes the model and muxing steps")
