#!/usr/bin/env python
"""bench.py -- stereo-reprojected frames/s at 1920x1080 (BASELINE.json metric, configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one 300-frame 1920x1080 synthetic clip per GPU
(decode + master-FOV scale + unproject + eye poses + re-projection + z-buffered splat + hole masks,
side-by-side output): ONE launch of the fused row kernel.  Frames shard across ranks with no data-path
collective (weak scaling: 300 frames per GPU; rank 0 broadcasts the per-frame constant block only).

  value     whole-job frames/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e       same clip through the public host API (StereoRerenderer.render_host): pinned host buffers,
            H2D + kernel + D2H inside the timed region
  roofline  algorithmic bytes (14 B/px: 6 in, 8 out) / mean launch time vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the NumPy oracle port of the reference path on a bounded sample (rank 0, N=1)
  result_codec  informational, N=1: the SBS frames of the step through the device FFV1 encoder (frames/s, bytes/frame);
            measured after the timed regions, never part of `value` / `e2e`

--impl reference times the reference's CPU path (NumPy oracle port; the reference itself needs
open3d + a Windows-only render()) on all host cores, a bounded sample per step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, HEIGHT, CLIP_FRAMES = 1920, 1080, 300
XFOV, MAX_DEPTH, IPD_MM, MASTER_XFOV = 60.0, 100, 63, 45.0
BYTES_PER_PX = 14  # depth u8x3 + colour u8x3 in; 2 x RGB u8x3 + 2 x mask u8 out
METRIC = "stereo-reprojected frames/sec at 1920x1080"
WORKLOAD = "configs[1]: 1920x1080 x 300-frame synthetic clip, stereo_rerender left/right warp + hole masks"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--frames", type=int, default=CLIP_FRAMES, help="frames per GPU per step")
    ap.add_argument("--distinct-frames", type=int, default=30, help="distinct synthetic frames generated per rank (tiled to --frames)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-frames", type=int, default=24, help="frames of the bounded CPU-baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-paths", action="store_true", help="skip the `paths` sections (convergence / posed stereo at 1080p, 4K novel view)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML in this process (a thread polling every 2 ms --
    the headline region is ~30 ms long, too short for a freshly started nvidia-smi to report from), nvidia-smi -lms as the
    fallback where pynvml is missing."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.samples, self.proc, self.gpu, self.source = [], None, gpu_index, None   # samples: (sm MHz, [reason, ...])
        self.sm_max, self._stop, self.thread = 0.0, threading.Event(), None

    def _nvml_handle(self):
        import pynvml

        pynvml.nvmlInit()
        try:   # CUDA ordinal -> NVML device through the UUID (CUDA_VISIBLE_DEVICES may renumber)
            import torch

            uuid = "GPU-" + str(torch.cuda.get_device_properties(self.gpu).uuid)
            try:
                return pynvml, pynvml.nvmlDeviceGetHandleByUUID(uuid)
            except TypeError:
                return pynvml, pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
        except Exception:  # noqa: BLE001
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.gpu)

    def _poll_nvml(self, pynvml, handle):
        masks = (("hw_slowdown", pynvml.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", pynvml.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", pynvml.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", pynvml.nvmlClocksEventReasonSwPowerCap))
        while not self._stop.is_set():
            try:
                mhz = float(pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM))
                try:
                    bits = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(handle))
                except Exception:  # noqa: BLE001 - older bindings
                    bits = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(handle))
                self.samples.append((mhz, [name for name, bit in masks if bits & bit]))
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.002)

    def __enter__(self):
        try:
            pynvml, handle = self._nvml_handle()
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
            self.source = "nvml"
            self.thread = threading.Thread(target=self._poll_nvml, args=(pynvml, handle), daemon=True)
            self.thread.start()
            return self
        except Exception:  # noqa: BLE001 - no NVML binding / no permission: ask nvidia-smi
            self.source = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.thread = threading.Thread(target=self._read_smi, daemon=True)
            self.thread.start()
            deadline = time.time() + 3.0   # nvidia-smi needs a few hundred ms before its first line
            while not self.samples and time.time() < deadline:
                time.sleep(0.01)
            self.samples.clear()
        except OSError:
            self.proc = None
        return self

    def _read_smi(self):
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.strip().split(",")]
            if len(parts) < 7:
                continue
            try:
                mhz = float(parts[0])
                self.sm_max = max(self.sm_max, float(parts[1]))
            except ValueError:
                continue
            self.samples.append((mhz, [n for n, v in zip(names, parts[3:7]) if v.lower().startswith("active")]))

    def __exit__(self, *exc):
        if self.source == "nvml":
            self._stop.set()
            self.thread.join(timeout=1)
        elif self.proc is not None:
            time.sleep(0.05)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm = sorted(m for m, _ in self.samples)
        reasons = sorted({r for _, rs in self.samples for r in rs})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.sm_max or None, "reasons": reasons, "samples": len(sm),
                "source": self.source}


# ------------------------------------------------------------------------------------------------
# CPU arms (the only places that touch oracle/)
# ------------------------------------------------------------------------------------------------
_CPU_FRAMES = {}  # pre-generated synthetic frames (inherited by forked workers): generation is never timed


def _prepare_cpu_frames(count: int):
    from metric_depth_video_toolbox_b200.synth import SyntheticClip

    clip = SyntheticClip(WIDTH, HEIGHT, CLIP_FRAMES)
    for k in range(count):
        if k not in _CPU_FRAMES:
            _CPU_FRAMES[k] = clip.frame(k)


def _oracle_stereo_frame(idx):
    from oracle import mdvt_oracle as orc

    depth_rgb, colour = _CPU_FRAMES[idx % len(_CPU_FRAMES)]
    t0 = time.perf_counter()
    sbs, mask, _ = orc.stereo_frame(depth_rgb, colour, XFOV, None, MAX_DEPTH, IPD_MM, MASTER_XFOV, infill_mask=True)
    return time.perf_counter() - t0, int(sbs[::97, ::89].sum()) + int(mask.sum() // 255)


def cpu_baseline(n_frames: int):
    """Oracle port, one core, frames already in RAM (generation is outside the timed part)."""
    _prepare_cpu_frames(n_frames)
    times = [_oracle_stereo_frame(i)[0] for i in range(n_frames)]
    total = sum(times)
    return {"value": n_frames / total, "unit": "frames/s", "cores": 1, "kind": "port",
            "sample": f"{n_frames} frames of the same 1920x1080 clip through oracle/mdvt_oracle.stereo_frame (NumPy float64, both eyes + masks), "
                      f"{total:.1f} s of CPU work"}


def _pin_worker(cpus, counter):
    """Pool initializer: every worker takes one host core of its own (no migration, no two workers on one core)."""
    with counter.get_lock():
        k = counter.value
        counter.value += 1
    try:
        os.sched_setaffinity(0, {cpus[k % len(cpus)]})
    except (AttributeError, OSError):
        pass


def workload_config(world: int, n_frames: int, distinct: int):
    """The `config` object of both arms (the reference arm runs bounded samples of the same workload)."""
    return {"workload": WORKLOAD, "width": WIDTH, "height": HEIGHT, "frames_per_gpu_per_step": n_frames,
            "distinct_frames_per_gpu": distinct, "xfov": XFOV, "max_depth": MAX_DEPTH, "pupillary_distance_mm": IPD_MM,
            "master_xfov": MASTER_XFOV, "sharding": f"frames, {world} rank(s), no data-path collective",
            "l2": f"inputs {2 * n_frames * HEIGHT * WIDTH * 3 / 1e9:.2f} GB + outputs {n_frames * HEIGHT * 2 * WIDTH * 4 / 1e9:.2f} GB per GPU per step "
                  f">> 126 MB L2 (no flush needed)"}


def run_reference(args):
    """The reference's CPU path (NumPy oracle port) on all host cores: frames are independent, so they are
    fanned over a process pool exactly like movie_2_3D.py --parallel fans scenes (movie_2_3D.py:422-452).
    One worker per core, pinned; frames stream through the pool with a window of tasks always outstanding (no barrier
    between steps: a "step" is every `4 x cores` completed frames); warm-up runs until two consecutive steps agree within
    5 %, at least --warmup steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from collections import deque

    cpus = sorted(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else list(range(os.cpu_count() or 1))
    cores = len(cpus)
    per_step = 4 * cores
    _prepare_cpu_frames(16)
    ctx = mp.get_context("fork")
    counter = ctx.Value("i", 0)
    with ctx.Pool(cores, initializer=_pin_worker, initargs=(cpus, counter)) as pool:
        inflight, submitted = deque(), 0

        def run_step():
            nonlocal submitted
            done = 0
            while done < per_step:
                while len(inflight) < 3 * cores:
                    inflight.append(pool.apply_async(_oracle_stereo_frame, (submitted,)))
                    submitted += 1
                inflight.popleft().get()
                done += 1

        warm, t_prev = [], time.perf_counter()
        while len(warm) < max(2, args.warmup) or (abs(warm[-1] - warm[-2]) > 0.05 * warm[-1] and len(warm) < args.warmup + 8):
            run_step()
            now = time.perf_counter()
            warm.append(now - t_prev)
            t_prev = now
        t0 = time.perf_counter()
        for _ in range(args.steps):
            run_step()
        elapsed = time.perf_counter() - t0
        for r in inflight:  # drain the window (outside the timed region)
            r.get()
    value = args.steps * per_step / elapsed
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": len(warm), "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(int(os.environ.get("WORLD_SIZE", "1")), args.frames, max(1, min(args.distinct_frames, args.frames))),
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": f"{per_step} frames per step ({args.steps} timed steps after {len(warm)} warm-up steps) of the same clip, streamed "
                                       f"through {cores} pinned single-threaded workers without a barrier between steps; NumPy float64 oracle port "
                                       f"(oracle/mdvt_oracle.stereo_frame, both eyes + masks)"},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic_per_frame():
    """DRAM bytes per frame of the fused kernel from the committed ncu capture (profiles/), if any, with the kernel name ncu
    reported and whether the capture was taken from the kernel source that is in the tree now (benchmarks/make_traffic_json.py
    stores the SHA-256 of the kernel's source files)."""
    path = os.path.join(ROOT, "profiles", "stereo_rows_traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as fh:
        traffic = json.load(fh)
    try:
        sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
        from make_traffic_json import source_sha256

        traffic["matches_current_source"] = traffic.get("kernel_source_sha256") == source_sha256()
    except Exception:  # noqa: BLE001
        traffic["matches_current_source"] = None
    return traffic


def vrows_traffic_per_frame():
    """DRAM bytes per 1080p frame of the virtual-source-row kernel from its committed ncu capture, or None (also None when the
    kernel's source has changed since the capture)."""
    path = os.path.join(ROOT, "profiles", "stereo_vrows_traffic.json")
    try:
        with open(path) as fh:
            t = json.load(fh)
        sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
        from make_traffic_json import VROWS_SOURCES, source_sha256

        return float(t["dram_bytes_per_frame"]) if t.get("kernel_source_sha256") == source_sha256(VROWS_SOURCES) else None
    except Exception:  # noqa: BLE001
        return None


def measure_paths(dev, peak: float, frames_1080: int = 32, frames_4k: int = 16, reps: int = 5):
    """The other kernels of the path at BASELINE.json's sizes, outside `value` (N = 1): convergence stereo (what
    movie_2_3D runs: --convergence_file), stereo with a pose file (--transformation_file), and configs[2] (3840x2160 novel
    view, camera (2, 2, -4) aimed at the frame's vertex centroid).  Each: CUDA-event time per frame on device-resident
    frames, the roofline fraction from SURVEY.md 8(d)'s algorithmic bytes (14 / 14 / 10 B/px), and the same frames through
    the host API with pinned host buffers (H2D + kernels + D2H inside the timed region)."""
    import math

    import numpy as np
    import torch

    from metric_depth_video_toolbox_b200.novel_view import NovelViewParams, NovelViewRenderer
    from metric_depth_video_toolbox_b200.stereo import StereoParams, StereoRerenderer
    from metric_depth_video_toolbox_b200.synth import SyntheticClip

    def clip(w, h, n, distinct=4):
        d, c = SyntheticClip(w, h, n).frames(0, distinct)
        reps_ = (n + distinct - 1) // distinct
        hd = torch.from_numpy(np.concatenate([d] * reps_)[:n]).pin_memory()
        hc = torch.from_numpy(np.concatenate([c] * reps_)[:n]).pin_memory()
        return hd, hc

    def timed(fn, n_frames):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (reps * n_frames)   # ms per frame

    def wall(fn, n_frames, steps=2):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3 / (steps * n_frames)

    def entry(ms_frame, e2e_ms_frame, w, h, bpp, kernel, h2d, d2h, note):
        achieved = bpp * w * h / (ms_frame / 1e3) / 1e9
        return {"ms_per_frame": ms_frame, "frames_per_s": 1e3 / ms_frame, "kernel": kernel, "workload": note,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "algorithmic_bytes_per_frame": bpp * w * h, "traffic": None},
                "e2e": {"value": 1e3 / e2e_ms_frame, "unit": "frames/s", "ms_per_frame": e2e_ms_frame, "h2d_bytes_per_frame": h2d,
                        "d2h_bytes_per_frame": d2h}}

    out = {}
    w, h, n = WIDTH, HEIGHT, frames_1080
    hd, hc = clip(w, h, n)
    d, c = hd.to(dev), hc.to(dev)
    sbs = torch.empty((n, h, 2 * w, 3), dtype=torch.uint8, device=dev)
    mask = torch.empty((n, h, 2 * w), dtype=torch.uint8, device=dev)
    h_sbs = torch.empty((n, h, 2 * w, 3), dtype=torch.uint8, pin_memory=True)
    h_mask = torch.empty((n, h, 2 * w), dtype=torch.uint8, pin_memory=True)
    poses = []
    for f in range(n):   # a small hand-held camera path: a few milliradians of rotation, centimetres of translation
        a, b = 0.01 * math.sin(0.21 * f), 0.004 * math.cos(0.13 * f)
        ry = np.array([[math.cos(a), 0, math.sin(a)], [0, 1, 0], [-math.sin(a), 0, math.cos(a)]])
        rx = np.array([[1, 0, 0], [0, math.cos(b), -math.sin(b)], [0, math.sin(b), math.cos(b)]])
        T = np.eye(4)
        T[:3, :3] = ry @ rx
        T[:3, 3] = (0.05 * math.sin(0.17 * f), -0.02 * math.cos(0.11 * f), 0.1 * math.sin(0.07 * f))
        poses.append(T)
    common = dict(xfov=XFOV, max_depth=MAX_DEPTH, pupillary_distance=IPD_MM, master_xfov=MASTER_XFOV, infill_mask=True)
    conv_list = [5.0 + 0.02 * f for f in range(n)]
    for key, extra, kernel, note in (
            ("convergence_1080p", dict(convergence_depths=conv_list), "mdvt::stereo_conv_vrows_kernel",
             f"{w}x{h} x {n} frames, stereo pair + hole masks with a per-frame convergence rotation (stereo_rerender --convergence_file, what movie_2_3D "
             f"runs), default dispatch: the virtual-source-row kernel (TMA-assembled rows, shared-memory z-buffer, no global planes)"),
            ("convergence_1080p_generic_loop", dict(convergence_depths=conv_list, conv_kernel="generic"),
             "mdvt::project_splat_kernel + mdvt::resolve_ckey_kernel", "the same frames through the two-lane generic loop (global 64-bit z-buffer planes; the bytes are identical)"),
            ("convergence_1080p_fused_row_kernel", dict(convergence_depths=conv_list, conv_kernel="rows"),
             "mdvt::stereo_conv_rows_kernel", "the same frames through round 1's target-row kernel (per-column source-row prediction, byte gathers)"),
            ("posed_1080p", dict(transformations=poses), "mdvt::project_splat_kernel + mdvt::resolve_ckey_kernel",
             f"{w}x{h} x {n} frames, stereo pair + hole masks with a per-frame 4x4 camera pose (stereo_rerender --transformation_file)")):
        rr = StereoRerenderer(StereoParams(w, h, **common, **extra), dev)
        ms = timed(lambda: rr.render_device(d, c, 0, sbs, mask), n)
        e2e_ms = wall(lambda: rr.render_host(hd, hc, h_sbs, h_mask), n)
        out[key] = entry(ms, e2e_ms, w, h, 14, kernel, 2 * w * h * 3, 2 * w * h * 4, note)
        out[key]["holes"] = float((mask == 255).float().mean().item())
        if key == "convergence_1080p":
            out[key]["roofline"]["traffic"] = vrows_traffic_per_frame()   # per frame, like algorithmic_bytes_per_frame
    del d, c, sbs, mask, h_sbs, h_mask, hd, hc
    w, h, n = 3840, 2160, frames_4k
    hd, hc = clip(w, h, n)
    d, c = hd.to(dev), hc.to(dev)
    nv = NovelViewRenderer(NovelViewParams(w, h, XFOV, None, MAX_DEPTH), dev)
    rgb = torch.empty((n, h, w, 3), dtype=torch.uint8, device=dev)
    msk = torch.empty((n, h, w), dtype=torch.uint8, device=dev)
    h_rgb = torch.empty((n, h, w, 3), dtype=torch.uint8, pin_memory=True)
    ms = timed(lambda: nv.render_device(d, c, 0, rgb, msk), n)
    e2e_ms = wall(lambda: nv.render_host(hd, hc, h_rgb), n)
    out["novel_view_4k"] = entry(ms, e2e_ms, w, h, 10, "mdvt::centroid_partial_kernel + mdvt::project_splat_kernel + mdvt::resolve_ckey_kernel",
                                 2 * w * h * 3, w * h * 3,
                                 f"configs[2]: {w}x{h} x {n} frames, 3d_view_depthfile --render, camera (2, 2, -4) aimed at the vertex centroid, white background")
    out["novel_view_4k"]["holes"] = float((msk == 255).float().mean().item())
    del d, c, rgb, msk, h_rgb, hd, hc, nv
    torch.cuda.empty_cache()
    out["movie_2_3D_steps_4_5_1080p"] = measure_movie_pipeline()
    return out


def measure_movie_pipeline(frames: int = 480):
    """BASELINE configs[4] end to end, files in -> files out: movie_2_3D step 4 (convergence list from the depth + focus-mask
    videos) and step 5 (stereo_rerender with the flags movie_2_3D passes, SBS + infill-mask videos) on a synthetic 1920x1080
    FFV1 clip in a temporary directory, through benchmarks/movie_e2e.py in a child process (wall clock, codecs included: the
    inputs are decoded and the results coded by the device FFV1 codec; green/black infill mask, i.e. without the host TELEA
    step of the normals-coded mask).  Informational: never costs the bench line."""
    import shutil
    import tempfile

    script = os.path.join(ROOT, "benchmarks", "movie_e2e.py")
    work = tempfile.mkdtemp(prefix="mdvt_bench_movie_")   # a few GB of FFV1 clips (the synthetic colour is incompressible): removed afterwards
    try:
        proc = subprocess.run([sys.executable, script, str(frames), "--green"], capture_output=True, text=True, timeout=600,
                              env=dict(os.environ, MDVT_E2E_DIR=os.path.join(work, "clip")))
        for line in reversed(proc.stdout.splitlines()):
            if line.startswith("{"):
                res = json.loads(line)
                res["frames"] = frames
                return res
        return {"error": (proc.stderr or proc.stdout)[-400:]}
    except Exception as exc:  # noqa: BLE001
        return {"error": f"{type(exc).__name__}: {exc}"}
    finally:
        shutil.rmtree(work, ignore_errors=True)


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from metric_depth_video_toolbox_b200 import ops
    from metric_depth_video_toolbox_b200.stereo import StereoParams, StereoRerenderer
    from metric_depth_video_toolbox_b200.synth import SyntheticClip

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: there is no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # host threads and the pinned buffers they first touch go next to this rank's GPU where the box exposes a topology
    from metric_depth_video_toolbox_b200 import sharding
    host_locality = sharding.bind_host_to_gpu(local_rank, world)
    saved_stdout = None
    if world > 1:
        # stdout carries exactly ONE JSON line (the contract): NCCL prints its version banner to fd 1 when the box sets
        # NCCL_DEBUG, so fd 1 points at stderr until the line is ready
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    n_frames = args.frames

    # ---- the clip shard of this rank (frames [rank*n, (rank+1)*n) of the whole job) -----------------------
    # `distinct` different synthetic frames are generated into a pinned host ring and tiled to n_frames on
    # the device: same bytes moved and same work per pixel as n_frames distinct frames, bounded host memory.
    distinct = max(1, min(args.distinct_frames, n_frames))
    clip = SyntheticClip(WIDTH, HEIGHT, world * n_frames)
    host_d = torch.empty((distinct, HEIGHT, WIDTH, 3), dtype=torch.uint8, pin_memory=True)
    host_c = torch.empty((distinct, HEIGHT, WIDTH, 3), dtype=torch.uint8, pin_memory=True)
    first = rank * n_frames
    for k in range(distinct):
        d, c = clip.frame(first + k)
        host_d[k] = torch.from_numpy(d)
        host_c[k] = torch.from_numpy(c)
    reps = (n_frames + distinct - 1) // distinct
    dev_d = host_d.to(dev).repeat(reps, 1, 1, 1)[:n_frames].contiguous()
    dev_c = host_c.to(dev).repeat(reps, 1, 1, 1)[:n_frames].contiguous()

    # ---- parameter block: built on rank 0, broadcast (the only collective of the path) --------------------
    params = StereoParams(WIDTH, HEIGHT, xfov=XFOV, max_depth=MAX_DEPTH, pupillary_distance=IPD_MM, master_xfov=MASTER_XFOV,
                          infill_mask=True)
    rr = StereoRerenderer(params, dev)
    consts = torch.from_numpy(rr.frame_constants(0, n_frames)).to(dev) if rank == 0 else torch.empty((1, 4), dtype=torch.float32, device=dev)
    if world > 1:
        dist.broadcast(consts, src=0)
    out_sbs = torch.empty((n_frames, HEIGHT, 2 * WIDTH, 3), dtype=torch.uint8, device=dev)
    out_mask = torch.empty((n_frames, HEIGHT, 2 * WIDTH), dtype=torch.uint8, device=dev)
    flags = ops.FLAG_BG_COLLIDE

    def step():
        ops.stereo_rows(dev_d, dev_c, consts, params.bg_rgb, (0, 0, 0), flags, out_sbs, out_mask)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step()
    sync_all()
    # ---- device-resident throughput ----------------------------------------------------------------------
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    with ClockSampler(local_rank) as clocks:
        sync_all()
        ev[0].record()
        for s in range(args.steps):
            step()
            ev[s + 1].record()
        sync_all()
    total_ms = ev[0].elapsed_time(ev[-1])
    launch_ms = [ev[s].elapsed_time(ev[s + 1]) for s in range(args.steps)]
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = world * n_frames * args.steps / (total_ms_max / 1e3)

    # ---- end to end through the host API -------------------------------------------------------------------
    # One e2e step = the same n_frames clip, streamed from the pinned host ring in `distinct`-frame segments
    # (like a reader filling a ring buffer) through StereoRerenderer.render_host: H2D, kernel and D2H are all
    # inside the timed region; bytes are counted from the tensors actually copied.
    e2e = None
    if not args.no_e2e:
        host_sbs = torch.empty((distinct, HEIGHT, 2 * WIDTH, 3), dtype=torch.uint8, pin_memory=True)
        host_mask = torch.empty((distinct, HEIGHT, 2 * WIDTH), dtype=torch.uint8, pin_memory=True)
        host_bits = torch.empty((distinct, HEIGHT, 2 * WIDTH // 8), dtype=torch.uint8, pin_memory=True)
        segments = [(f0, min(distinct, n_frames - f0)) for f0 in range(0, n_frames, distinct)]

        def e2e_run(mask_format):
            out_mask_host = host_bits if mask_format == "bits" else host_mask

            def e2e_step():
                for f0, cnt in segments:
                    rr.render_host(host_d[:cnt], host_c[:cnt], host_sbs[:cnt], out_mask_host[:cnt], start_frame=f0, mask_format=mask_format)

            e2e_step()  # warm-up: allocates the device staging buffers
            sync_all()
            t0 = time.perf_counter()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.e2e_steps):
                e2e_step()
            e1.record()
            sync_all()
            wall_ms = (time.perf_counter() - t0) * 1e3
            t = torch.tensor([max(e0.elapsed_time(e1), wall_ms)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return world * n_frames * args.e2e_steps / (float(t.item()) / 1e3), int(out_mask_host[::17].to(torch.int64).sum().item())

        # the API's default (one byte per mask pixel) and its packed form (one bit per pixel: same information, the mask's
        # share of the device-to-host bytes drops from 25 % to 4 %); the packed form is the headline, both are reported
        u8_value, u8_checksum = e2e_run("u8")
        bits_value, bits_checksum = e2e_run("bits")
        per_frame_in = 2 * HEIGHT * WIDTH * 3
        per_frame_out = HEIGHT * 2 * WIDTH * 3 + HEIGHT * 2 * WIDTH // 8
        e2e = {"value": bits_value, "unit": "frames/s", "h2d_bytes_per_step": int(per_frame_in * n_frames * world),
               "d2h_bytes_per_step": int(per_frame_out * n_frames * world), "steps": args.e2e_steps,
               "api": "StereoRerenderer.render_host(mask_format='bits') (pinned host ring, 3 staging buffers with a stream each: chunked H2D/kernel/D2H; hole mask shipped as "
                      "one bit per pixel, ops.unpack_mask_bits restores the u8 plane)",
               "mask_checksum": bits_checksum,
               "u8_mask": {"value": u8_value, "unit": "frames/s", "d2h_bytes_per_step": int(HEIGHT * 2 * WIDTH * 4 * n_frames * world),
                           "api": "StereoRerenderer.render_host() (default: one byte per mask pixel)", "mask_checksum": u8_checksum},
               "host": host_locality}

    # ---- informational: the result codec on the device (N=1 only; outside every timed region above) ----------
    # The side-by-side frames the timed kernel just wrote, through mdvt_ffv1_encode_frames (FFV1 as the reference's result
    # writers produce it, stereo_rerender.py:420-442,941).  Never fails the bench line: errors are reported in the key.
    result_codec = None
    if world == 1 and not args.no_e2e:
        try:
            from metric_depth_video_toolbox_b200 import ffv1_gpu

            batch = min(64, n_frames)    # one thread per slice with 1 KB of coder state each: 64 frames is the sweet spot (3 840 x 1 080, film-like content: 4.30 k frames/s at 64, 3.36 k at 128 and 256 -- the states of 128 k slices no longer fit the L2; profiles/r02_ffv1_gpu_bench_batch_sweep.jsonl)
            result_codec = {"what": "FFV1 v3 entropy coding of the SBS result on the device (mdvt_ffv1_encode_frames), not part of `value`",
                            "frame": f"{2 * WIDTH}x{HEIGHT}", "batch": batch}

            def time_encoder(frames_dev, model, frames_per_launch):
                enc = ffv1_gpu.Ffv1Encoder(2 * WIDTH, HEIGHT, dev, max_frames=frames_per_launch, context_model=model)
                enc.encode_device(frames_dev[:frames_per_launch])
                torch.cuda.synchronize()
                c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                c0.record()
                for _ in range(3):
                    _, offsets = enc.encode_device(frames_dev[:frames_per_launch])
                c1.record()
                torch.cuda.synchronize()
                ms = c0.elapsed_time(c1) / 3
                return {"frames_per_s": 1e3 * frames_per_launch / ms, "ms_per_frame": ms / frames_per_launch, "batch": frames_per_launch,
                        "slices_per_frame": enc.per_frame, "bytes_per_frame": int(offsets[-1].item()) / frames_per_launch}

            # model 2 keeps its 224 bytes of state per slice in shared memory and wants more slices in flight (256 frames per launch)
            for model, key, per_launch in ((0, "libavcodec_tables_666_contexts", batch), (1, "small_tables_63_contexts", batch),
                                           (2, "tiny_tables_14_contexts", min(256, n_frames))):
                result_codec[key] = time_encoder(out_sbs, model, per_launch)
            # the same coder on film-like content (a blurred texture that pans, sensor noise in the low bits, a flat patch and a
            # black band: benchmarks/ffv1_gpu_bench.frames_like) -- the rendered synthetic clip above is i.i.d. colour noise, which
            # no lossless coder compresses (13.8 MB of packets per 12.4 MB frame) and which takes the long escape codes everywhere
            sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
            from ffv1_gpu_bench import frames_like

            film = torch.from_numpy(frames_like(2 * WIDTH, HEIGHT, 256)).to(dev)
            result_codec["film_like_content_63_contexts"] = time_encoder(film, 1, 64)
            result_codec["film_like_content_14_contexts"] = time_encoder(film, 2, 256)
            del film
        except Exception as exc:  # noqa: BLE001 - informational leg only
            result_codec = {"error": f"{type(exc).__name__}: {exc}"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_kind = measured_peak()
    paths = None
    if world == 1 and not args.no_paths:
        dev_d = dev_c = out_sbs = out_mask = None   # the 8.7 GB of the headline step make room for the 4K frames
        torch.cuda.empty_cache()
        try:
            paths = measure_paths(dev, peak)
        except Exception as exc:  # noqa: BLE001 - informational leg: never costs the bench line
            paths = {"error": f"{type(exc).__name__}: {exc}"}
    mean_launch_ms = sum(launch_ms) / len(launch_ms)
    algorithmic = BYTES_PER_PX * WIDTH * HEIGHT * n_frames
    achieved = algorithmic / (mean_launch_ms / 1e3) / 1e9
    traffic = ncu_traffic_per_frame()
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": (traffic["dram_bytes_per_frame"] * n_frames) if traffic else None,
                "kernel": (traffic or {}).get("kernel") or "mdvt::stereo_rows_w32_kernel<1,true,160,4>", "algorithmic_bytes_per_launch": algorithmic,
                "launch_ms": mean_launch_ms, "peak_source": peak_kind,
                "traffic_source": traffic.get("source") if traffic else None,
                "traffic_capture_matches_kernel_source": traffic.get("matches_current_source") if traffic else None}
    line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(world, n_frames, distinct),
            "roofline": roofline, "e2e": e2e, "gpu_launches": args.steps, "clocks": clocks.summary()}
    if paths is not None:
        line["paths"] = paths
    if result_codec is not None:
        line["result_codec"] = result_codec
    if not args.no_cpu and world == 1:
        line["cpu_baseline"] = cpu_baseline(args.cpu_frames)
    if saved_stdout is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    print(json.dumps(line), flush=True)
    if world > 1:
        sys.stdout.flush()
        os.dup2(2, 1)   # teardown messages, if any, do not follow the JSON line on stdout
        dist.destroy_process_group()


def main():
    args = parse_args()
    # the CPU legs are one NumPy process per core: BLAS / OpenMP pools of their own would only oversubscribe the cores
    if args.impl == "reference":
        for var in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
            os.environ.setdefault(var, "1")
        if os.environ.get("MDVT_BENCH_MALLOC") != "1" and int(os.environ.get("RANK", "0")) == 0:
            # NumPy's 16 MB float64 temporaries otherwise go through mmap / munmap on every operation (measured here: 30 s of
            # system time per 3 min of user time, -17 % throughput); glibc reads these at start-up, so the arm re-executes itself
            env = dict(os.environ, MDVT_BENCH_MALLOC="1", MALLOC_MMAP_THRESHOLD_="33554432", MALLOC_TRIM_THRESHOLD_="2000000000",
                       MALLOC_TOP_PAD_="268435456")
            sys.stdout.flush()
            os.execve(sys.executable, [sys.executable] + sys.argv, env)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
