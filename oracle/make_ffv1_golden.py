"""Writes tests/golden/ffv1_opencv.npz: a 5-frame 32x24 clip coded by this image's OpenCV (cv2.VideoWriter, fourcc FFV1 --
the writer behind every result video of the reference: stereo_rerender.py:420-442,941; depth_frames_helper.py:125-161) with
its CodecPrivate and packets, so that the FFV1 oracle and the slice coder stay pinned on libavcodec's output even where
another OpenCV / libavcodec build writes different stream parameters.

    python oracle/make_ffv1_golden.py        (run where OpenCV has an FFV1 encoder; records cv2 / avcodec versions)
"""
import os
import re
import sys
import tempfile

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from metric_depth_video_toolbox_b200 import mkv_join  # noqa: E402

W, H, N = 32, 24, 5


def frames():
    rng = np.random.default_rng(20261017)
    out = [rng.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(N)]
    out[1] = cv2.GaussianBlur(out[1], (0, 0), 2)         # smooth: small residuals
    out[2][:] = out[2][:1, :1]                           # flat: run mode
    out[3] = np.where(out[3] < 128, 0, 255).astype(np.uint8)   # extreme residuals: escape codes
    return np.stack(out)


def main():
    f = frames()
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "cv.mkv")
        wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"FFV1"), 24.0, (W, H))
        assert wr.isOpened()
        for x in f:
            wr.write(x)
        wr.release()
        pk = mkv_join.MkvPackets(path)
        packets = [bytes(pk.payload(k)) for k in range(N)]
        keys = [bool(p[2]) for p in pk.packets]
        config = bytes(pk.codec_private())
    m = re.search(r"avcodec\s+YES \(([^)]+)\)", cv2.getBuildInformation())
    out = os.path.join(ROOT, "tests", "golden", "ffv1_opencv.npz")
    np.savez_compressed(out, frames_bgr=f, config=np.frombuffer(config, np.uint8), keys=np.array(keys),
                        packet_sizes=np.array([len(p) for p in packets]), packets=np.frombuffer(b"".join(packets), np.uint8),
                        versions=np.array([cv2.__version__, m.group(1) if m else "?"]))
    print(out, os.path.getsize(out), "bytes; keys", keys, "packet sizes", [len(p) for p in packets])


if __name__ == "__main__":
    main()
