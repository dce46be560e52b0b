"""Extract the command-line surface (flag, type, default, action, required) of the reference scripts into
tests/golden/cli_flags.json, by parsing their `parser.add_argument(...)` calls with `ast` (the scripts
keep everything under `if __name__ == '__main__'`, so they cannot be imported).

    python oracle/make_cli_golden.py        # needs /root/reference (or MDVT_REFERENCE_ROOT)

TEST INFRASTRUCTURE ONLY.  Holds flag metadata, no reference source.
"""
from __future__ import annotations

import ast
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("MDVT_REFERENCE_ROOT", "/root/reference")
SCRIPTS = ["stereo_rerender.py", "3d_view_depthfile.py", "convert_metric_depth_video_to_other_format.py", "find_convergence_depth.py"]


def flags_of(path):
    tree = ast.parse(open(path).read())
    out = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr == "add_argument":
            name = ast.literal_eval(node.args[0])
            spec = {"type": None, "default": None, "action": None, "required": False}
            for kw in node.keywords:
                if kw.arg == "type":
                    spec["type"] = kw.value.id
                elif kw.arg in ("default", "action", "required"):
                    spec[kw.arg] = ast.literal_eval(kw.value)
            out[name] = spec
    return out


if __name__ == "__main__":
    golden = {s: flags_of(os.path.join(REF, s)) for s in SCRIPTS}
    dst = os.path.join(ROOT, "tests", "golden", "cli_flags.json")
    with open(dst, "w") as fh:
        json.dump(golden, fh, indent=1, sort_keys=True)
    print(dst, {k: len(v) for k, v in golden.items()})
