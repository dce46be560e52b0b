"""Deterministic synthetic RGB-encoded depth + colour clips.

Data generator only: it is neither on the measured path nor a CPU fallback for it.  Frames
come out exactly as the reference scripts see them after ``cv2.cvtColor(BGR2RGB)``
(stereo_rerender.py:493): (H, W, 3) uint8, RGB order, 16-bit wire format (R = G = high byte,
B = low byte of the top 16 bits of the 32-bit code, depth_frames_helper.py:48-61).
"""
from __future__ import annotations

import numpy as np

FULL_SCALE = 255 ** 4


class SyntheticClip:
    """Background ramp 8 + 2 sin(2 pi i / H) metres plus `n_rects` axis-aligned foreground
    rectangles at U(1, 4) m, each U(5 %, 25 %) of the frame per dimension, sliding right by
    `step_px` pixels per frame (step edges => real disocclusion holes).  `zero_fraction` of
    the pixels are forced to code 0 (exercises the z <= near skip rule).  Colour is i.i.d.
    uniform bytes so that a mis-mapped pixel is detectable."""

    def __init__(self, width, height, n_frames, seed=1234, max_depth=100, n_rects=6, step_px=2,
                 zero_fraction=0.0):
        self.width, self.height, self.n_frames = int(width), int(height), int(n_frames)
        self.seed, self.max_depth = int(seed), max_depth
        self.step_px, self.zero_fraction = int(step_px), float(zero_fraction)
        rng = np.random.default_rng(self.seed)
        self.rects = []
        for _ in range(n_rects):
            rw = max(1, int(rng.uniform(0.05, 0.25) * self.width))
            rh = max(1, int(rng.uniform(0.05, 0.25) * self.height))
            x0 = int(rng.integers(0, max(1, self.width - rw)))
            y0 = int(rng.integers(0, max(1, self.height - rh)))
            self.rects.append((x0, y0, rw, rh, float(rng.uniform(1.0, 4.0))))
        rows = np.arange(self.height, dtype=np.float64)
        self._background = (8.0 + 2.0 * np.sin(2.0 * np.pi * rows / self.height)).astype(np.float32)

    def depth_metres(self, frame_idx: int) -> np.ndarray:
        d = np.repeat(self._background[:, None], self.width, axis=1)
        # far rectangles first so nearer ones overwrite them
        for x0, y0, rw, rh, z in sorted(self.rects, key=lambda r: -r[4]):
            xs = (x0 + self.step_px * frame_idx) % self.width
            cols = (xs + np.arange(rw)) % self.width
            d[y0:y0 + rh, cols] = np.float32(z)
        return d

    def frame(self, frame_idx: int):
        """-> (depth_rgb u8 (H,W,3), colour_rgb u8 (H,W,3))."""
        rng = np.random.default_rng(self.seed + 1 + frame_idx)
        depth = self.depth_metres(frame_idx)
        codes = ((FULL_SCALE / float(self.max_depth)) * np.clip(depth, 0.0, self.max_depth).astype(np.float64)).astype(np.uint32)
        if self.zero_fraction > 0:
            codes[rng.random(codes.shape) < self.zero_fraction] = 0
        hi = (codes >> 24).astype(np.uint8)
        lo = ((codes >> 16) & 0xFF).astype(np.uint8)
        depth_rgb = np.stack((hi, hi, lo), axis=-1)
        colour = rng.integers(0, 256, size=(self.height, self.width, 3), dtype=np.uint8)
        return depth_rgb, colour

    def frames(self, start=0, stop=None):
        stop = self.n_frames if stop is None else stop
        n = max(0, stop - start)
        depth = np.empty((n, self.height, self.width, 3), dtype=np.uint8)
        colour = np.empty_like(depth)
        for k in range(n):
            depth[k], colour[k] = self.frame(start + k)
        return depth, colour
