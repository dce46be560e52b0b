"""Import the *real* reference modules from a read-only checkout, when one is reachable.

TEST INFRASTRUCTURE ONLY (see oracle/mdvt_oracle.py).  Used by oracle/make_golden.py to
produce tests/golden/*.npz and by tests/test_oracle_vs_reference.py, which skips itself when
the checkout is absent (it is absent on the GPU box).

`depth_map_tools.py` / `stereo_rerender.py` import open3d / PyOpenGL / glfw at module scope;
none is installed here, and none is needed by the pure NumPy / cv2 functions we call, so empty
stand-in modules are registered first.  No reference source is copied: inline script code is
read from the checkout at run time and exec()'d by line range (`exec_lines`).
"""
from __future__ import annotations

import importlib
import os
import sys
import textwrap
import types

REFERENCE_ROOT = os.environ.get("MDVT_REFERENCE_ROOT", "/root/reference")
_STUBS = ("open3d", "OpenGL", "OpenGL.GL", "OpenGL.GL.shaders", "glfw")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "depth_map_tools.py"))


def load(name: str):
    """Import reference module `name` (e.g. 'depth_frames_helper') under an alias so that it can
    never shadow this repo's drop-in module of the same name."""
    if not available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    alias = f"_mdvt_reference_.{name}"
    if alias in sys.modules:
        return sys.modules[alias]
    for stub in _STUBS:
        if stub not in sys.modules:
            if stub == "open3d":
                from . import fake_o3d

                sys.modules[stub] = fake_o3d.module()  # enough of Open3D to RUN the mesh-building NumPy code
            else:
                sys.modules[stub] = types.ModuleType(stub)
    shaders = sys.modules["OpenGL.GL.shaders"]
    for attr in ("compileProgram", "compileShader"):
        if not hasattr(shaders, attr):
            setattr(shaders, attr, None)
    # reference modules import each other by bare name; resolve those inside the checkout only
    saved_path = list(sys.path)
    saved_mods = {k: sys.modules.pop(k) for k in ("depth_frames_helper", "depth_map_tools", "stereo_rerender")
                  if k in sys.modules}
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        spec = importlib.util.spec_from_file_location(alias, os.path.join(REFERENCE_ROOT, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[alias] = mod
        spec.loader.exec_module(mod)
    finally:
        sys.path[:] = saved_path
        for k in ("depth_frames_helper", "depth_map_tools", "stereo_rerender"):
            sys.modules.pop(k, None)
        sys.modules.update(saved_mods)
    return mod


def exec_lines(filename: str, first: int, last: int, namespace: dict) -> dict:
    """exec() lines first..last (1-based, inclusive) of a reference script in `namespace`.
    This is how the inline decoders and the point painter of the scripts' __main__ blocks are
    run without importing (or copying) them."""
    with open(os.path.join(REFERENCE_ROOT, filename), "r", encoding="utf-8") as fh:
        lines = fh.readlines()[first - 1:last]
    exec(compile(textwrap.dedent("".join(lines)), f"{filename}:{first}-{last}", "exec"), namespace)
    return namespace
