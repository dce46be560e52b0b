"""Golden vectors for the host helpers of movie_2_3D.py (scene splitting / planning, time codes, the csv loader, the flag
table), produced by calling the reference's own functions:

    python oracle/make_movie_golden.py        # needs /root/reference (or MDVT_REFERENCE_ROOT)

TEST INFRASTRUCTURE ONLY.  Holds inputs and the reference's outputs, no reference source.
"""
from __future__ import annotations

import ast
import copy
import importlib.util
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("MDVT_REFERENCE_ROOT", "/root/reference")


def load_reference():
    sys.path.insert(0, REF)   # movie_2_3D imports depth_frames_helper
    spec = importlib.util.spec_from_file_location("_ref_movie_2_3D", os.path.join(REF, "movie_2_3D.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def flags_of(path):
    out = {}
    for node in ast.walk(ast.parse(open(path).read())):
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr == "add_argument":
            name = ast.literal_eval(node.args[0])
            spec = {"type": None, "default": None, "action": None, "required": False}
            for kw in node.keywords:
                if kw.arg == "type":
                    spec["type"] = kw.value.id
                elif kw.arg in ("default", "action", "required"):
                    try:
                        spec[kw.arg] = ast.literal_eval(kw.value)
                    except ValueError:
                        spec[kw.arg] = "expr:" + ast.unparse(kw.value)
            out[name] = spec
    return out


SCENES = [
    {"Scene Number": "1", "Start Frame": "1", "Start Timecode": "00:00:00.000", "Start Time (seconds)": "0.000", "End Frame": "120",
     "End Timecode": "00:00:05.005", "End Time (seconds)": "5.005", "Length (frames)": "120", "Length (timecode)": "00:00:05.005",
     "Length (seconds)": "5.005"},
    {"Scene Number": "2", "Start Frame": "121", "Start Timecode": "00:00:05.005", "Start Time (seconds)": "5.005", "End Frame": "3500",
     "End Timecode": "00:02:25.979", "End Time (seconds)": "145.979", "Length (frames)": "3380", "Length (timecode)": "00:02:20.974",
     "Length (seconds)": "140.974", "Infill": "No", "xfov": "63.5"},
    {"Scene Number": "3", "Start Frame": "3501", "Start Timecode": "00:02:25.979", "Start Time (seconds)": "145.979", "End Frame": "3501",
     "End Timecode": "00:02:25.979", "End Time (seconds)": "145.979", "Length (frames)": "1", "Length (timecode)": "00:00:00.000",
     "Length (seconds)": "0.000", "Convergence": "No"},
    {"Scene Number": "4", "Start Frame": "3502", "Start Timecode": "00:02:26.021", "Start Time (seconds)": "146.021", "End Frame": "5001",
     "End Timecode": "00:03:28.542", "End Time (seconds)": "208.542", "Length (frames)": "1500", "Length (timecode)": "00:01:02.521",
     "Length (seconds)": "62.521"},
]
CSV_TEXT = "Timecode List:,00:00:05.005,00:02:25.979\n" + ",".join(SCENES[0].keys()) + "\n" + \
    "\n".join(",".join(s[k] for k in SCENES[0].keys()) for s in SCENES) + "\n"

if __name__ == "__main__":
    ref = load_reference()
    golden = {"flags": flags_of(os.path.join(REF, "movie_2_3D.py")),
              "timecodes": {repr(x): ref._seconds_to_timecode(x) for x in (0.0, 0.0004, 0.0005, 1.0015, 59.9996, 61.5, 3599.9995, 3725.042, 86399.999)},
              "scenes": SCENES, "split": {}, "plan": {}, "csv_text": CSV_TEXT, "csv": {}}
    for max_frames in (1500, 1000, 7):
        golden["split"][str(max_frames)] = ref.split_scenes(copy.deepcopy(SCENES), max_scene_frames=max_frames)
    missing = os.path.join(tempfile.gettempdir(), "mdvt_no_such_dir_for_golden")
    for end_scene in (-1, 2):
        planned = ref.plan_scene_files(ref.split_scenes(copy.deepcopy(SCENES), 1500), missing, end_scene)
        golden["plan"][str(end_scene)] = [{k: (v.replace(missing, "<OUT>") if isinstance(v, str) else v) for k, v in s.items()} for s in planned]
    with tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False, newline="") as fh:
        fh.write(CSV_TEXT)
    for max_frames in (1500, 900):
        golden["csv"][str(max_frames)] = ref.load_and_split_scenes(fh.name, ",", max_frames)
    os.unlink(fh.name)
    dst = os.path.join(ROOT, "tests", "golden", "movie_2_3D.json")
    with open(dst, "w") as out:
        json.dump(golden, out, indent=1, sort_keys=True)
    print(dst, {k: (len(v) if hasattr(v, "__len__") else v) for k, v in golden.items()})
