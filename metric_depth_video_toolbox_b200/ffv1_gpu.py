"""FFV1 result videos coded on the device (`mdvt_ffv1_encode_frames`, csrc/mdvt_ffv1.cu) and muxed here.

The reference writes every result through `cv2.VideoWriter(..., fourcc "FFV1", ...)` (stereo_rerender.py:420-442,941;
depth_frames_helper.py:125-161; 3d_view_depthfile.py:118-127): ~0.45 core-seconds of entropy coding per 3840x1080 frame,
which caps the scripts at a few frames/s whatever renders the frames.  Here the rendered frames never
leave the device uncompressed: one thread codes one FFV1 slice (up to 1024 per frame, a batch of frames per launch), the
slices are compacted into packets on the device, and only the packets cross PCIe; `mkv_join.StreamWriter` puts them in a
Matroska file whose header and track description are the ones OpenCV/FFmpeg write for that size and rate, with the codec
configuration record replaced by the one of this stream.  The stream is standard FFV1 version 3 (same coder, colour
transform, quant tables and CRC as OpenCV's files; every frame a key frame, more slices) and decodes bit-identically.

There is no CPU implementation: without the library / a device the encoder raises.
"""
from __future__ import annotations

import ctypes as C
import os
import tempfile
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib, mkv_join

HEADER_STRIDE = 16


def slice_grid(width: int, height: int, target_pixels: int = 4096) -> Tuple[int, int]:
    """Slices per row / column: as many as FFV1 allows (1024) while a slice keeps about target_pixels pixels -- the
    per-slice coder state needs something to adapt on -- and is roughly square (fewest border samples)."""
    total = max(1, min(1024, (width * height) // max(1, target_pixels)))
    nv = max(1, min(height, total, int(round((total * height / width) ** 0.5))))
    nh = max(1, min(width, total // nv))
    return nh, nv


def stream_setup(width: int, height: int, nh: int, nv: int, alpha: bool = False, context_model: int = 0):
    """(configuration record bytes, slice headers (S, 16) uint8, header lengths (S,) int32) -- host arrays.
    context_model 0: libavcodec's quant tables (666 contexts), 1: the 5-level table (63 contexts), 2: the 3-level table (14 contexts:
    the device encoder keeps the coder states in shared memory; ~1 % larger files than model 1)."""
    lib = _lib.load()
    config = (C.c_uint8 * 64)()
    n = C.c_int(0)
    headers = np.zeros((nh * nv, HEADER_STRIDE), np.uint8)
    lens = np.zeros(nh * nv, np.int32)
    _lib.check(lib.mdvt_ffv1_stream_setup(width, height, nh, nv, int(alpha), int(context_model), C.addressof(config), 64, C.byref(n),
                                          headers.ctypes.data, lens.ctypes.data))
    return bytes(config[:n.value]), headers, lens


class Ffv1Encoder:
    """Device buffers + stream constants for frames of one size; `encode` turns a batch of device frames into packets."""

    def __init__(self, width: int, height: int, device, max_frames: int = 8, slices: Optional[Tuple[int, int]] = None,
                 alpha: bool = False, context_model: int = 0):
        self.lib = _lib.load()
        self.width, self.height, self.alpha, self.context_model = int(width), int(height), bool(alpha), int(context_model)
        self.nh, self.nv = slices if slices is not None else slice_grid(width, height)
        self.per_frame = self.nh * self.nv
        self.max_frames = int(max_frames)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.MdvtError(-4, "the FFV1 encoder runs on a CUDA device only")
        self.config, headers, lens = stream_setup(width, height, self.nh, self.nv, alpha, self.context_model)
        self.capacity = int(self.lib.mdvt_ffv1_slice_capacity(width, height, self.nh, self.nv, int(alpha)))
        state_bytes = int(self.lib.mdvt_ffv1_state_bytes(self.max_frames, self.nh, self.nv, int(alpha), self.context_model))
        if self.capacity <= 0 or state_bytes < 0:
            raise ValueError(f"bad FFV1 stream parameters {width}x{height}, {self.nh}x{self.nv} slices")
        n_slices = self.max_frames * self.per_frame
        dev = self.device
        self.headers = torch.from_numpy(headers).to(dev)
        self.header_len = torch.from_numpy(lens).to(dev)
        self.states = torch.empty(state_bytes, dtype=torch.uint8, device=dev)
        self.slices = torch.empty(n_slices * self.capacity, dtype=torch.uint8, device=dev)
        self.packed = torch.empty(n_slices * self.capacity, dtype=torch.uint8, device=dev)
        self.sizes = torch.empty(n_slices, dtype=torch.int32, device=dev)
        self.offsets = torch.empty(n_slices + 1, dtype=torch.int64, device=dev)
        self._host = torch.empty(0, dtype=torch.uint8).pin_memory()

    def encode_device(self, frames: torch.Tensor, rgb: bool = True):
        """frames: (n <= max_frames, H, W, 3) uint8 on the device.  Enqueues the encode on the current stream and
        returns (packed device bytes, offsets device int64 (n * S + 1,)) -- views of this encoder's buffers."""
        if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[3] != 3 or not frames.is_cuda:
            raise TypeError("expected a CUDA uint8 tensor (n, H, W, 3)")
        n, h, w = int(frames.shape[0]), int(frames.shape[1]), int(frames.shape[2])
        if (h, w) != (self.height, self.width) or n > self.max_frames:
            raise ValueError(f"encoder built for <= {self.max_frames} frames of {self.width}x{self.height}, got {n} of {w}x{h}")
        if frames.stride(3) != 1 or frames.stride(2) != 3:
            frames = frames.contiguous()
        stream = torch.cuda.current_stream(frames.device).cuda_stream
        _lib.check(self.lib.mdvt_ffv1_encode_frames(
            frames.data_ptr(), frames.stride(0), frames.stride(1), n, w, h, self.nh, self.nv, int(self.alpha), self.context_model,
            0 if rgb else 1,
            self.headers.data_ptr(), self.header_len.data_ptr(), self.states.data_ptr(), self.slices.data_ptr(), self.capacity,
            self.sizes.data_ptr(), self.offsets.data_ptr(), self.packed.data_ptr(), stream))
        return self.packed, self.offsets[: n * self.per_frame + 1]

    def encode(self, frames: torch.Tensor, rgb: bool = True):
        """-> list of n packets (bytes), one per frame.  Two device-to-host copies: the offsets, then the used bytes."""
        n = int(frames.shape[0])
        if n == 0:
            return []
        packed, offsets = self.encode_device(frames, rgb)
        bounds = offsets[:: self.per_frame].cpu().numpy()   # synchronises with the encode
        total = int(bounds[-1])
        if self._host.numel() < total:
            self._host = torch.empty(max(total, 2 * self._host.numel()), dtype=torch.uint8).pin_memory()
        self._host[:total].copy_(packed[:total])
        buf = self._host.numpy()
        return [buf[int(bounds[k]):int(bounds[k + 1])].tobytes() for k in range(n)]


class Ffv1Decoder:
    """The mirror image of Ffv1Encoder: packets of a stream written with this library's parameters -> frames on the device,
    one device thread per slice (`mdvt_ffv1_decode_frames`)."""

    def __init__(self, width: int, height: int, device, max_frames: int = 8, slices: Optional[Tuple[int, int]] = None,
                 alpha: bool = False, context_model: int = 0):
        self.lib = _lib.load()
        self.width, self.height, self.alpha, self.context_model = int(width), int(height), bool(alpha), int(context_model)
        self.nh, self.nv = slices if slices is not None else slice_grid(width, height)
        self.per_frame = self.nh * self.nv
        self.max_frames = int(max_frames)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.MdvtError(-4, "the FFV1 decoder runs on a CUDA device only")
        self.config, headers, lens = stream_setup(width, height, self.nh, self.nv, alpha, self.context_model)
        state_bytes = int(self.lib.mdvt_ffv1_state_bytes(self.max_frames, self.nh, self.nv, int(alpha), self.context_model))
        if state_bytes < 0:
            raise ValueError(f"bad FFV1 stream parameters {width}x{height}, {self.nh}x{self.nv} slices")
        dev = self.device
        self.headers = torch.from_numpy(headers).to(dev)
        self.header_len = torch.from_numpy(lens).to(dev)
        self.states = torch.empty(state_bytes, dtype=torch.uint8, device=dev)
        self.slice_offsets = torch.empty(self.max_frames * self.per_frame, dtype=torch.int64, device=dev)
        self.status = torch.empty(self.max_frames, dtype=torch.int32, device=dev)
        self._host = torch.empty(0, dtype=torch.uint8).pin_memory()
        self._dev = torch.empty(0, dtype=torch.uint8, device=dev)

    @classmethod
    def for_config(cls, config: bytes, width: int, height: int, device, max_frames: int = 8):
        """A decoder for the stream whose configuration record (Matroska CodecPrivate) is `config`; raises MdvtError when
        the stream was not written with this library's parameters."""
        lib = _lib.load()
        nh, nv, alpha, model = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _lib.check(lib.mdvt_ffv1_parse_config(config, len(config), width, height, C.byref(nh), C.byref(nv), C.byref(alpha), C.byref(model)))
        return cls(width, height, device, max_frames, (nh.value, nv.value), bool(alpha.value), model.value)

    def decode(self, packets, rgb: bool = True, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """packets: a sequence of n <= max_frames packets (bytes-like).  Returns (n, H, W, 3) uint8 on the device (RGB
        order by default); raises MdvtError naming the first frame that could not be decoded."""
        n = len(packets)
        if n > self.max_frames:
            raise ValueError(f"decoder built for <= {self.max_frames} frames per call, got {n}")
        if out is None:
            out = torch.empty((n, self.height, self.width, 3), dtype=torch.uint8, device=self.device)
        elif tuple(out.shape) != (n, self.height, self.width, 3) or out.dtype != torch.uint8 or not out.is_cuda or not out.is_contiguous():
            raise TypeError("`out` must be a contiguous CUDA uint8 tensor (n, H, W, 3)")
        if n == 0:
            return out
        bounds = np.zeros(n + 1, np.int64)
        np.cumsum([len(p) for p in packets], out=bounds[1:])
        total = int(bounds[-1])
        if self._host.numel() < total:
            self._host = torch.empty(max(total, 2 * self._host.numel()), dtype=torch.uint8).pin_memory()
            self._dev = torch.empty(self._host.numel(), dtype=torch.uint8, device=self.device)
        host = self._host.numpy()
        for k, p in enumerate(packets):
            host[int(bounds[k]):int(bounds[k + 1])] = np.frombuffer(p, np.uint8)
        self._dev[:total].copy_(self._host[:total], non_blocking=True)
        offsets = torch.from_numpy(bounds).to(self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.mdvt_ffv1_decode_frames(
            self._dev.data_ptr(), offsets.data_ptr(), n, self.width, self.height, self.nh, self.nv, int(self.alpha), self.context_model,
            0 if rgb else 1,
            self.headers.data_ptr(), self.header_len.data_ptr(), self.states.data_ptr(), self.slice_offsets.data_ptr(), out.data_ptr(),
            out.stride(0), out.stride(1), self.status.data_ptr(), stream))
        status = self.status[:n].cpu().numpy()   # synchronises: the pinned staging buffer is free again
        if (status != 0).any():
            k = int(np.nonzero(status)[0][0])
            reason = {-2: "slice sizes do not add up", -3: "foreign slice header (not a key frame of this library's stream)",
                      -4: "inconsistent slice size", -5: "bit stream overrun", -6: "slice CRC mismatch (damaged data)"}.get(int(status[k]), "unknown")
            raise _lib.MdvtError(int(status[k]), f"FFV1 packet {k} of {n}: {reason}")
        return out


class GpuFfv1Reader:
    """Frames of an .mkv written by GpuFfv1Writer, decoded on the device: iterating yields (n, H, W, 3) uint8 CUDA tensors
    (RGB order by default) of up to `batch` frames.  Files with other FFV1 parameters (e.g. cv2.VideoWriter's) raise
    MdvtError on open: they stay with cv2.VideoCapture on the host, as in the reference."""

    def __init__(self, path: str, device=None, batch: int = 8, rgb: bool = True, start: int = 0, stop: Optional[int] = None):
        from . import video_io

        self.path, self.rgb, self.batch = path, rgb, max(1, batch)
        self.width, self.height, self.fps, _ = video_io.video_info(path)
        self._pk = mkv_join.MkvPackets(path)
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.dec = Ffv1Decoder.for_config(self._pk.codec_private(), self.width, self.height, self.device, max_frames=self.batch)
        self.frames = len(self._pk.packets)
        if not all(key for _, _, key in self._pk.packets):   # cv2.VideoWriter's 2 x 2 + alpha record equals this library's
            raise _lib.MdvtError(-2, f"{path}: the stream carries coder state across frames (non-key frames); only all-key-frame "
                                     "streams are decoded on the device")
        self.start, self.stop = max(0, start), self.frames if stop is None else min(stop, self.frames)

    def __len__(self):
        return max(0, self.stop - self.start)

    def __iter__(self):
        for a in range(self.start, self.stop, self.batch):
            b = min(a + self.batch, self.stop)
            yield self.dec.decode([self._pk.payload(k) for k in range(a, b)], rgb=self.rgb)

    def read_all(self) -> torch.Tensor:
        chunks = list(self)
        return torch.cat(chunks) if chunks else torch.empty((0, self.height, self.width, 3), dtype=torch.uint8, device=self.device)

    def close(self):
        self._pk.close()


_TEMPLATES: dict = {}


def container_template(width: int, height: int, fps: float):
    """(EBML header, Tracks payload) of the file OpenCV/FFmpeg write for an FFV1 video of this size and rate (made once
    per size and rate: cv2.VideoWriter codes a whole frame for it)."""
    key = (int(width), int(height), float(fps))
    if key not in _TEMPLATES:
        _TEMPLATES[key] = _make_container_template(*key)
    return _TEMPLATES[key]


def _make_container_template(width: int, height: int, fps: float):
    import cv2

    fd, path = tempfile.mkstemp(suffix=".mkv", prefix="mdvt_ffv1_template_")
    os.close(fd)
    try:
        w = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"FFV1"), fps, (width, height))
        if not w.isOpened():
            raise RuntimeError("cannot open an FFV1 writer for the container template")
        w.write(np.zeros((height, width, 3), np.uint8))
        w.release()
        pk = mkv_join.MkvPackets(path)
        out = pk.ebml_header, pk.tracks
        pk.close()
        return out
    finally:
        if os.path.exists(path):
            os.remove(path)


def default_context_model() -> int:
    """The writers' context model: 1 (63 contexts) unless MDVT_FFV1_CONTEXT_MODEL names 0, 1 or 2."""
    v = os.environ.get("MDVT_FFV1_CONTEXT_MODEL", "").strip()
    if not v:
        return 1
    if v not in ("0", "1", "2"):
        raise ValueError(f"MDVT_FFV1_CONTEXT_MODEL must be 0, 1 or 2, not {v!r}")
    return int(v)


class GpuFfv1Writer:
    """`cv2.VideoWriter(path, FFV1, fps, size)` for frames that live on the device.  write() takes (n, H, W, 3) uint8
    tensors (device; host tensors / arrays are uploaded before it returns, so the caller may recycle its buffer), RGB by
    default; close() finishes the file.  Encoding, the packet download and the muxing run on a worker thread, at most
    `depth` write() calls behind the caller."""

    def __init__(self, path: str, fps: float, size: Tuple[int, int], device=None, batch: int = 8,
                 slices: Optional[Tuple[int, int]] = None, alpha: bool = False, join_on_close: bool = True, depth: int = 2,
                 context_model: Optional[int] = None):
        """context_model 1 (the default, or what MDVT_FFV1_CONTEXT_MODEL says): the 63-context quant table -- 1 KB instead of
        10.6 KB of coder state per slice thread, ~40 % more frames/s on the B200 and files within -0.5 .. +1 % of the
        666-context ones (the tables travel in the configuration record: any FFV1 decoder reads either); 0: libavcodec's
        own tables; 2: the 14-context table whose coder states the kernels keep in shared memory (~1.9 x the encoder
        throughput of model 1 for ~1 % larger files).
        join_on_close=False: `path` is one rank's segment of a torchrun job; close() leaves `<path>.plan.json` next to
        it in video_io.ParallelWriter's format, and rank 0 stitches the segments at packet level (video_io.join_plans)."""
        import queue
        import threading

        self.path, self.fps, self.size = path, fps, (int(size[0]), int(size[1]))
        self.join_on_close = join_on_close
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.enc = Ffv1Encoder(self.size[0], self.size[1], self.device, max_frames=batch, slices=slices, alpha=alpha,
                               context_model=default_context_model() if context_model is None else context_model)
        header, tracks = container_template(self.size[0], self.size[1], fps)
        self._mux = mkv_join.StreamWriter(path, header, mkv_join.replace_codec_private(tracks, self.enc.config), fps)
        self.frames = 0
        self.bytes = 0
        self._error: Optional[BaseException] = None
        self._queue: "queue.Queue" = queue.Queue(maxsize=max(1, depth))
        self._worker = threading.Thread(target=self._run, name="mdvt-ffv1-writer", daemon=True)
        self._worker.start()

    def _run(self):
        torch.cuda.set_device(self.device)
        # a stream of its own: on the (legacy) default stream the encode launches of this thread would queue behind -- and
        # hold up -- the renderer's kernels and the other writers' encodes, and the coder, a latency-bound kernel with one
        # thread per slice, is exactly the kind of work that should run underneath them
        stream = torch.cuda.Stream(device=self.device)
        while True:
            item = self._queue.get()
            if item is None:
                return
            if self._error is not None:
                continue   # keep draining so the producer never blocks
            t, rgb = item
            try:
                with torch.cuda.stream(stream):
                    for a in range(0, int(t.shape[0]), self.enc.max_frames):
                        for pkt in self.enc.encode(t[a:a + self.enc.max_frames], rgb):
                            self._mux.add(pkt, True)
                            self.bytes += len(pkt)
            except BaseException as e:   # surfaced by the next write() / close()
                self._error = e

    def _raise_pending(self):
        if self._error is not None:
            e, self._error = self._error, None
            raise e

    def write(self, frames, rgb: bool = True):
        self._raise_pending()
        t = frames if isinstance(frames, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(frames))
        if t.dim() == 3:
            t = t[None]
        if t.dtype != torch.uint8 or t.dim() != 4 or t.shape[3] != 3:
            raise TypeError("expected uint8 frames (n, H, W, 3)")
        if tuple(t.shape[1:3]) != (self.size[1], self.size[0]):
            raise ValueError(f"frames are {t.shape[2]}x{t.shape[1]}, writer expects {self.size[0]}x{self.size[1]}")
        if t.shape[0] == 0:
            return
        # an own device copy, complete before returning (blocking upload / stream-ordered clone + synchronise)
        if t.is_cuda:
            t = t.clone()
            torch.cuda.current_stream(t.device).synchronize()
        else:
            t = t.to(self.device)
        self.frames += int(t.shape[0])
        self._queue.put((t, rgb))

    def _stop(self):
        if self._worker.is_alive():
            self._queue.put(None)
            self._worker.join()

    def close(self) -> int:
        self._stop()
        if self._error is not None:
            self._mux.abort()
            self._raise_pending()
        n = self._mux.close()
        if not self.join_on_close:
            import json

            with open(self.path + ".plan.json", "w") as fh:
                json.dump({"fps": self.fps, "plan": [(self.path, n)] if n else []}, fh)
            if n == 0 and os.path.exists(self.path):   # a rank without frames leaves no segment behind
                os.remove(self.path)
        return n

    def abort(self):
        self._stop()
        self._mux.abort()
