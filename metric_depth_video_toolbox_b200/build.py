"""Build libmdvt_b200.so in-tree with plain nvcc for sm_100a (cross-compiles without a GPU).

    python -m metric_depth_video_toolbox_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB_PATH = os.path.join(PKG_DIR, "libmdvt_b200.so")

SOURCES = ["mdvt_abi.cu", "mdvt_pointwise.cu", "mdvt_splat.cu", "mdvt_reduce.cu", "mdvt_export.cu", "mdvt_edges.cu", "mdvt_remap.cu", "mdvt_stereo_rows.cu", "mdvt_stereo_conv.cu", "mdvt_stereo_vrows.cu", "mdvt_ffv1.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--fmad=false",            # no FMA contraction anywhere: every float op rounds once (parity contract)
    "-Xcompiler", "-fPIC,-O2,-Wall,-fvisibility=hidden",
    "--shared", "-cudart", "shared",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libmdvt_b200.so cannot be built")
    return exe


def _inputs():
    files = [os.path.join(CSRC, s) for s in SOURCES]
    files += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    files += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    return files


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(f) > built for f in _inputs())


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every source to an object file (in parallel: the sources are independent) and link the shared library."""
    if not force and not needs_build():
        return LIB_PATH
    import tempfile
    from concurrent.futures import ThreadPoolExecutor

    nvcc = _nvcc()
    compile_flags = [f for f in NVCC_FLAGS if f != "--shared"]
    with tempfile.TemporaryDirectory(prefix="mdvt_b200_build_") as tmp:
        def compile_one(src: str):
            obj = os.path.join(tmp, os.path.splitext(src)[0] + ".o")
            cmd = [nvcc, *compile_flags, "-I", INCLUDE, "-I", CSRC] + (["-Xptxas", "-v"] if verbose else []) + \
                ["-c", os.path.join(CSRC, src), "-o", obj]
            return obj, subprocess.run(cmd, capture_output=True, text=True)

        with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as pool:
            results = list(pool.map(compile_one, SOURCES))
        failed = [r for _, r in results if r.returncode != 0]
        if verbose or failed:
            for _, r in results:
                sys.stderr.write(r.stdout + r.stderr)
        if failed:
            raise RuntimeError("nvcc failed building libmdvt_b200.so")
        tmp_lib = LIB_PATH + ".linking"
        link = subprocess.run([nvcc, *NVCC_FLAGS, *[obj for obj, _ in results], "-o", tmp_lib], capture_output=True, text=True)
        if verbose or link.returncode != 0:
            sys.stderr.write(link.stdout + link.stderr)
        if link.returncode != 0:
            raise RuntimeError("nvcc failed linking libmdvt_b200.so")
        os.replace(tmp_lib, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
