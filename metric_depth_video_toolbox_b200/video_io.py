"""Host-side clip I/O for the script front ends: OpenCV readers / writers on background threads that fill
and drain pinned chunk buffers, so FFV1 decode/encode (the real end-to-end limiter) overlaps
the H2D -> kernel -> D2H pipeline.  Container handling is the reference's (cv2.VideoCapture /
cv2.VideoWriter, FFV1 or avc1); nothing per-pixel besides the BGR<->RGB byte swap happens here.
"""
from __future__ import annotations

import queue
import threading
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch


def _cv2():
    import cv2

    return cv2


def video_info(path: str) -> Tuple[int, int, float, int]:
    """(width, height, fps, frame_count) as the scripts read them (stereo_rerender.py:375-377)."""
    cv2 = _cv2()
    cap = cv2.VideoCapture(path)
    try:
        return (int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)), int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT)), cap.get(cv2.CAP_PROP_FPS),
                int(cap.get(cv2.CAP_PROP_FRAME_COUNT)))
    finally:
        cap.release()


class ChunkReader:
    """Reads frames [start, stop) of one or more same-length videos in lock step, converts BGR -> RGB and
    yields chunks of up to `chunk` frames as pinned uint8 tensors (n, H, W, 3) -- one tensor per video.
    A background thread stays `depth` chunks ahead.  A short video ends the iteration for all of them
    (the scripts stop at the first failed read, stereo_rerender.py:489-503)."""

    def __init__(self, paths: Sequence[Optional[str]], start: int = 0, stop: Optional[int] = None, chunk: int = 8, depth: int = 3,
                 pin: bool = True, grey: Sequence[bool] = ()):
        cv2 = _cv2()
        self.paths = list(paths)
        self.caps = [None if p is None else cv2.VideoCapture(p) for p in self.paths]
        self.grey = list(grey) + [False] * (len(self.paths) - len(grey))
        first = next(c for c in self.caps if c is not None)
        self.width, self.height = int(first.get(cv2.CAP_PROP_FRAME_WIDTH)), int(first.get(cv2.CAP_PROP_FRAME_HEIGHT))
        total = min(int(c.get(cv2.CAP_PROP_FRAME_COUNT)) for c in self.caps if c is not None)
        self.start, self.stop = start, total if stop is None else min(stop, total)
        if start > 0:
            for c in self.caps:
                if c is not None:
                    c.set(cv2.CAP_PROP_POS_FRAMES, start)  # FFV1 / intra-only: exact
        self.chunk, self.pin = max(1, chunk), pin and torch.cuda.is_available()
        self._q: "queue.Queue" = queue.Queue(maxsize=max(1, depth))
        self._free: "queue.Queue" = queue.Queue()
        for _ in range(depth + 2):
            self._free.put(self._alloc())
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def _alloc(self) -> List[Optional[torch.Tensor]]:
        bufs = []
        for c, g in zip(self.caps, self.grey):
            shape = (self.chunk, self.height, self.width) + (() if g else (3,))
            bufs.append(None if c is None else torch.empty(shape, dtype=torch.uint8, pin_memory=self.pin))
        return bufs

    def _run(self):
        cv2 = _cv2()
        pos = self.start
        try:
            while pos < self.stop:
                bufs = self._free.get()
                n = 0
                while n < self.chunk and pos < self.stop:
                    ok_all = True
                    for cap, buf, g in zip(self.caps, bufs, self.grey):
                        if cap is None:
                            continue
                        ok, frame = cap.read()
                        if not ok:
                            ok_all = False
                            break
                        cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY if g else cv2.COLOR_BGR2RGB, dst=buf[n].numpy())
                    if not ok_all:
                        pos = self.stop
                        break
                    n += 1
                    pos += 1
                if n:
                    self._q.put((n, bufs))
        except BaseException as exc:  # surfaced on the consumer side
            self._q.put(exc)
        finally:
            self._q.put(None)
            for c in self.caps:
                if c is not None:
                    c.release()

    def __iter__(self):
        while True:
            item = self._q.get()
            if item is None:
                return
            if isinstance(item, BaseException):
                raise item
            n, bufs = item
            yield n, [None if b is None else b[:n] for b in bufs]
            self._free.put(bufs)  # the consumer is done with the previous chunk when it asks for the next


class ChunkWriter:
    """cv2.VideoWriter on a background thread.  `write(frames_rgb)` takes a (n, H, W, 3) uint8 host tensor /
    array in RGB order (or BGR with rgb=False, as the depth encoder produces) and returns immediately; the
    bytes are copied before it returns, so the caller may reuse its buffer."""

    def __init__(self, path: str, fourcc: str, fps: float, size: Tuple[int, int], depth: int = 4):
        cv2 = _cv2()
        self.path, self.size = path, size
        self.writer = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*fourcc), fps, size)
        if not self.writer.isOpened():
            raise RuntimeError(f"cannot open a {fourcc} writer for {path}")
        self.frames = 0
        self._q: "queue.Queue" = queue.Queue(maxsize=depth)
        self._err: Optional[BaseException] = None
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def _run(self):
        cv2 = _cv2()
        try:
            while True:
                item = self._q.get()
                if item is None:
                    return
                frames, rgb = item
                for f in frames:
                    self.writer.write(cv2.cvtColor(f, cv2.COLOR_RGB2BGR) if rgb else f)
        except BaseException as exc:
            self._err = exc

    def write(self, frames, rgb: bool = True):
        if self._err is not None:
            raise self._err
        arr = frames.numpy() if isinstance(frames, torch.Tensor) else np.asarray(frames)
        if arr.shape[1:3] != (self.size[1], self.size[0]):
            raise ValueError(f"frames are {arr.shape[2]}x{arr.shape[1]}, writer expects {self.size[0]}x{self.size[1]}")
        self.frames += arr.shape[0]
        self._q.put((arr.copy(), rgb))

    def close(self):
        self._q.put(None)
        self._thread.join()
        self.writer.release()
        if self._err is not None:
            raise self._err


def write_clip(path: str, frames_rgb, fps: float = 24.0, fourcc: str = "FFV1", rgb: bool = True):
    """Small helper for tests / synthetic inputs: (n, H, W, 3) uint8 -> video file."""
    frames_rgb = np.asarray(frames_rgb)
    w = ChunkWriter(path, fourcc, fps, (frames_rgb.shape[2], frames_rgb.shape[1]))
    w.write(frames_rgb, rgb=rgb)
    w.close()


def read_clip(path: str, rgb: bool = True) -> np.ndarray:
    cv2 = _cv2()
    cap = cv2.VideoCapture(path)
    frames = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        frames.append(cv2.cvtColor(f, cv2.COLOR_BGR2RGB) if rgb else f)
    cap.release()
    return np.stack(frames) if frames else np.zeros((0, 0, 0, 3), dtype=np.uint8)
