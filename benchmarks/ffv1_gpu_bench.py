"""Device FFV1 encoder: frames/s and file size against OpenCV's single-threaded writer on the same frames.

    python benchmarks/ffv1_gpu_bench.py [--width 3840 --height 1080 --frames 16 --batch 16 --reps 3]

Content: the side-by-side result of the stereo path on the synthetic clip when the package can render it, else blurred
noise with rectangles.  Prints one JSON line per measurement."""
import argparse
import json
import os
import sys
import tempfile
import time

import cv2
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metric_depth_video_toolbox_b200 import ffv1_gpu  # noqa: E402


def frames_like(w, h, n, seed=0):
    rng = np.random.default_rng(seed)
    base = cv2.GaussianBlur(rng.integers(0, 256, (h, w, 3), dtype=np.uint8), (0, 0), 2.0)
    out = []
    for k in range(n):
        f = np.roll(base, 3 * k, axis=1).copy()
        f += rng.integers(0, 4, f.shape, dtype=np.uint8)          # sensor-like noise in the low bits
        f[h // 4: h // 2, (w // 8 + 5 * k) % (w // 2): (w // 8 + 5 * k) % (w // 2) + w // 6] = (20, 200, 90)
        f[:, w // 2 - 8: w // 2] = 0                              # a disocclusion-like black band
        out.append(f)
    return np.stack(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--cv_frames", type=int, default=2)
    ap.add_argument("--context_model", type=int, default=0, help="0: libavcodec's 666 contexts, 1: 63 contexts, 2: 14 contexts (coder states in shared memory)")
    ap.add_argument("--decode", action="store_true", help="also time Ffv1Decoder.decode on the packets (H2D of the packets + kernels + status read)")
    ap.add_argument("--encode_only", action="store_true", help="skip the cv2.VideoWriter and file legs")
    ap.add_argument("--grids", default="auto,32x32,16x16", help="slice grids to time: auto or NHxNV, comma separated")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    frames = frames_like(a.width, a.height, a.frames)
    d = torch.from_numpy(frames).to(dev)
    for grid in a.grids.split(","):
        slices = None if grid == "auto" else tuple(int(v) for v in grid.split("x"))
        enc = ffv1_gpu.Ffv1Encoder(a.width, a.height, dev, max_frames=a.batch, slices=slices, context_model=a.context_model)
        enc.encode_device(d[: a.batch])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            for s in range(0, a.frames, a.batch):
                enc.encode_device(d[s: s + a.batch])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        t0 = time.perf_counter()
        nbytes = 0
        for s in range(0, a.frames, a.batch):
            nbytes += sum(len(p) for p in enc.encode(d[s: s + a.batch]))
        host_s = time.perf_counter() - t0
        print(json.dumps({"what": "device FFV1 encode", "size": [a.width, a.height], "slices": [enc.nh, enc.nv], "context_model": a.context_model, "frames": a.frames,
                          "batch": a.batch, "device_ms_per_frame": ms / a.frames, "device_frames_per_s": 1000.0 * a.frames / ms,
                          "with_d2h_frames_per_s": a.frames / host_s, "bytes_per_frame": nbytes / a.frames,
                          "bits_per_pixel": 8.0 * nbytes / a.frames / (a.width * a.height)}), flush=True)
        if a.decode:
            packets = enc.encode(d[: a.batch])
            dec = ffv1_gpu.Ffv1Decoder.for_config(enc.config, a.width, a.height, dev, max_frames=a.batch)
            out = torch.empty((len(packets), a.height, a.width, 3), dtype=torch.uint8, device=dev)
            dec.decode(packets, out=out)          # warm-up: staging buffers
            t0 = time.perf_counter()
            for _ in range(a.reps):
                dec.decode(packets, out=out)
            dt = (time.perf_counter() - t0) / a.reps
            print(json.dumps({"what": "device FFV1 decode (packets in host memory -> frames on the device)", "slices": [dec.nh, dec.nv],
                              "context_model": dec.context_model, "batch": len(packets), "frames_per_s": len(packets) / dt,
                              "ms_per_frame": 1e3 * dt / len(packets), "identical": bool(torch.equal(out, d[: a.batch]))}), flush=True)
            del dec, out
        del enc
    if a.encode_only:
        return
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "cv.mkv")
        t0 = time.perf_counter()
        w = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"FFV1"), 24.0, (a.width, a.height))
        for f in frames[: a.cv_frames]:
            w.write(f)
        w.release()
        cv_s = time.perf_counter() - t0
        print(json.dumps({"what": "cv2.VideoWriter FFV1 (one thread)", "frames": a.cv_frames, "frames_per_s": a.cv_frames / cv_s,
                          "bytes_per_frame": os.path.getsize(path) / a.cv_frames}), flush=True)
        gpath = os.path.join(tmp, "gpu.mkv")
        t0 = time.perf_counter()
        gw = ffv1_gpu.GpuFfv1Writer(gpath, 24.0, (a.width, a.height), device=dev, batch=a.batch, context_model=a.context_model)
        t1 = time.perf_counter()
        gw.write(d)
        gw.close()
        t2 = time.perf_counter()
        cap = cv2.VideoCapture(gpath)
        ok_all, k = True, 0
        while k < min(a.frames, 3):
            ok, got = cap.read()
            ok_all &= bool(ok) and np.array_equal(cv2.cvtColor(got, cv2.COLOR_BGR2RGB), frames[k])
            k += 1
        print(json.dumps({"what": "GpuFfv1Writer file (device frames -> .mkv)", "frames": a.frames, "open_s": t1 - t0,
                          "frames_per_s": a.frames / (t2 - t1), "bytes_per_frame": os.path.getsize(gpath) / a.frames,
                          "first_frames_decode_identically": ok_all}), flush=True)


if __name__ == "__main__":
    main()
