"""Host-side clip I/O for the script front ends: OpenCV readers / writers on background threads that fill
and drain pinned chunk buffers, so FFV1 decode/encode (the real end-to-end limiter) overlaps
the H2D -> kernel -> D2H pipeline.  Container handling is the reference's (cv2.VideoCapture /
cv2.VideoWriter, FFV1 or avc1); nothing per-pixel besides the BGR<->RGB byte swap happens here.
"""
from __future__ import annotations

import json
import os
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


def _is_ffv1(cap) -> bool:
    cv2 = _cv2()
    cc = int(cap.get(cv2.CAP_PROP_FOURCC))
    return bytes((cc >> (8 * k)) & 0xFF for k in range(4)).upper() == b"FFV1"


def default_decoders(world_size: int = 1) -> int:
    """Decode threads per input video.  1: every input has its own sequential decoder.  More (MDVT_READER_THREADS) gives
    every decoder its own chunks.  With seeking decoders that was measured counter-productive (OpenCV's frame-exact seek
    lands on the key frame at or before (target - 16) and decodes forward from there, ~22 frames per seek with a GOP of
    12: 3.5 -> 1.2 frames/s on two 4K inputs); Matroska inputs are now served without seeking (the chunk's packets are
    copied into a small temporary file, _DecodeWorker).  One FFV1 decoder already runs up to four slice threads, so more
    decoders only pay on hosts with more cores than 4 x inputs: on 8 cores one noisy 1080p input alone went 16 -> 22
    frames/s with four such decoders, two inputs in lock step 16.9 -> 14.7 (cores already saturated); not yet measured
    on the 16-core GPU box, hence still opt-in."""
    env = os.environ.get("MDVT_READER_THREADS")
    return max(1, int(env)) if env else 1


class _DecodeWorker(threading.Thread):
    """One cv2.VideoCapture on its own thread: fills (buffer, first frame, count) requests.  A request that does not
    continue where the previous one ended is served, when the input is a Matroska file `mkv_join` can read, from a small
    temporary file holding just the packets from the key frame at or before `first` to the end of the request (a packet
    copy: no seek, no decoding of frames nobody asked for beyond that GOP); otherwise by seeking."""

    def __init__(self, path: str, grey: bool, cut: bool = False):
        super().__init__(daemon=True)
        self.path, self.grey, self.cut = path, grey, cut
        self.tasks: "queue.Queue" = queue.Queue()
        self.start()

    def submit(self, buf, first: int, count: int) -> dict:
        ticket = {"done": threading.Event(), "n": 0, "err": None}
        self.tasks.put((buf, first, count, ticket))
        return ticket

    def _read(self, cap, out, count: int, code) -> int:
        cv2 = _cv2()
        n = 0
        while n < count:
            ok, frame = cap.read()
            if not ok:
                break
            cv2.cvtColor(frame, code, dst=out[n])
            n += 1
        return n

    def _read_cut(self, packets, fps: float, out, first: int, count: int, code) -> int:
        """Frames [first, first + count) through a temporary file of their packets."""
        import tempfile

        from . import mkv_join

        cv2 = _cv2()
        total = len(packets.packets)
        if first >= total:
            return 0
        key = first
        while key > 0 and not packets.packets[key][2]:
            key -= 1
        stop = min(total, first + count)
        fd, tmp = tempfile.mkstemp(suffix=".mkv", prefix="mdvt_cut_")
        os.close(fd)
        try:
            mkv_join.write_stream(tmp, packets.ebml_header, packets.tracks,
                                  ((packets.payload(k), packets.packets[k][2]) for k in range(key, stop)), stop - key, fps)
            cap = cv2.VideoCapture(tmp)
            for _ in range(first - key):
                if not cap.grab():
                    break
            n = self._read(cap, out, stop - first, code)
            cap.release()
            return n
        finally:
            if os.path.exists(tmp):
                os.remove(tmp)

    def run(self):
        cv2 = _cv2()
        cap = cv2.VideoCapture(self.path)
        fps = cap.get(cv2.CAP_PROP_FPS) or 24.0
        packets = None
        if self.cut:
            try:
                from . import mkv_join

                packets = mkv_join.MkvPackets(self.path)
            except Exception:   # not a file mkv_join reads: seek instead
                packets = None
        pos = 0
        code = cv2.COLOR_BGR2GRAY if self.grey else cv2.COLOR_BGR2RGB
        while True:
            item = self.tasks.get()
            if item is None:
                cap.release()
                if packets is not None:
                    packets.close()
                return
            buf, first, count, ticket = item
            try:
                out = buf.numpy()
                if first != pos and packets is not None:
                    ticket["n"] = self._read_cut(packets, fps, out, first, count, code)
                else:
                    if first != pos:
                        cap.set(cv2.CAP_PROP_POS_FRAMES, first)  # block starts are key frames of FFV1 / intra-only sources: exact
                        pos = first
                    n = self._read(cap, out, count, code)
                    pos += n
                    ticket["n"] = n
            except BaseException as exc:
                ticket["err"] = exc
            ticket["done"].set()


def _can_cut(path: str) -> bool:
    """True when `mkv_join` can take the file apart packet by packet (what _DecodeWorker needs to serve a frame range
    without seeking)."""
    from . import mkv_join

    try:
        pk = mkv_join.MkvPackets(path)
        ok = len(pk.packets) > 0 and pk.packets[0][2]
        pk.close()
        return bool(ok)
    except Exception:
        return False


class ChunkReader:
    """Reads frames [start, stop) of one or more same-length videos in lock step, converts BGR -> RGB and
    yields chunks of up to `chunk` frames as pinned uint8 tensors (n, H, W, 3) -- one tensor per video.
    Every video has its own decode thread(s); background threads stay `depth` chunks ahead.  `decoders` > 1 decodes
    several chunks of the same video at once (FFV1 sources with GOP-aligned chunks only: each decoder seeks to the
    chunk's first frame, a key frame).  A short video ends the iteration for all of them (the scripts stop at the first
    failed read, stereo_rerender.py:489-503)."""

    def __init__(self, paths: Sequence[Optional[str]], start: int = 0, stop: Optional[int] = None, chunk: int = 8, depth: int = 3,
                 pin: bool = True, grey: Sequence[bool] = (), decoders: int = 1):
        cv2 = _cv2()
        self.paths = list(paths)
        self.grey = list(grey) + [False] * (len(self.paths) - len(grey))
        probes = [None if p is None else cv2.VideoCapture(p) for p in self.paths]
        first = next(c for c in probes if c is not None)
        self.width, self.height = int(first.get(cv2.CAP_PROP_FRAME_WIDTH)), int(first.get(cv2.CAP_PROP_FRAME_HEIGHT))
        total = min(int(c.get(cv2.CAP_PROP_FRAME_COUNT)) for c in probes if c is not None)
        all_ffv1 = all(_is_ffv1(c) for c in probes if c is not None)
        for c in probes:
            if c is not None:
                c.release()
        self.start, self.stop = start, total if stop is None else min(stop, total)
        self.chunk, self.pin = max(1, chunk), pin and torch.cuda.is_available()
        # parallel decode: every decoder is handed the packets of its chunk (any chunk size / start) when the containers can be
        # taken apart; else it needs exact seeks: FFV1 written with OpenCV's GOP, chunks that start on key frames
        cut = decoders > 1 and all_ffv1 and all(_can_cut(p) for p in self.paths if p is not None)
        self.decoders = max(1, decoders) if (cut or (all_ffv1 and self.chunk % GOP == 0 and start % GOP == 0)) else 1
        self._workers = [None if p is None else [_DecodeWorker(p, g, cut) for _ in range(self.decoders)]
                         for p, g in zip(self.paths, self.grey)]
        self._q: "queue.Queue" = queue.Queue(maxsize=max(1, depth))
        self._free: "queue.Queue" = queue.Queue()
        for _ in range(depth + 1 + self.decoders):   # pinned: 300 MB per 12-frame 4K chunk, keep the ring short
            self._free.put(self._alloc())
        self._tickets: "queue.Queue" = queue.Queue(maxsize=max(1, depth) + self.decoders)
        self._stop, self._tickets_done, self._finished = threading.Event(), threading.Event(), threading.Event()
        self._q_end: list = []   # the exception the collector ended with, if any
        self._threads = [threading.Thread(target=self._dispatch, daemon=True), threading.Thread(target=self._collect, daemon=True)]
        for t in self._threads:
            t.start()

    def _alloc(self) -> List[Optional[torch.Tensor]]:
        bufs = []
        for p, g in zip(self.paths, self.grey):
            shape = (self.chunk, self.height, self.width) + (() if g else (3,))
            bufs.append(None if p is None else torch.empty(shape, dtype=torch.uint8, pin_memory=self.pin))
        return bufs

    # blocking queue operations poll, so that close() (early `break`, an exception in the consumer, interpreter exit) can
    # always get the background threads out of them
    def _get(self, q: "queue.Queue"):
        while not self._stop.is_set():
            try:
                return q.get(timeout=0.05)
            except queue.Empty:
                continue
        return None

    def _put(self, q: "queue.Queue", item) -> bool:
        while not self._stop.is_set():
            try:
                q.put(item, timeout=0.05)
                return True
            except queue.Full:
                continue
        return False

    def _dispatch(self):
        c = 0
        pos = self.start
        try:
            while pos < self.stop and not self._stop.is_set():
                bufs = self._get(self._free)
                if bufs is None:
                    break
                count = min(self.chunk, self.stop - pos)
                tickets = [None if w is None else w[c % self.decoders].submit(b, pos, count) for w, b in zip(self._workers, bufs)]
                if not self._put(self._tickets, (count, bufs, tickets)):
                    break
                pos += count
                c += 1
        finally:
            self._tickets_done.set()

    def _collect(self):
        end = None
        try:
            while True:
                try:
                    item = self._tickets.get(timeout=0.05)
                except queue.Empty:
                    if self._stop.is_set() or self._tickets_done.is_set():
                        break
                    continue
                count, bufs, tickets = item
                n = count
                for t in tickets:
                    if t is None:
                        continue
                    t["done"].wait()
                    if t["err"] is not None:
                        raise t["err"]
                    n = min(n, t["n"])
                if n and not self._put(self._q, (n, bufs)):
                    break
                if n < count:   # a video ended early: stop here for all of them
                    break
        except BaseException as exc:  # surfaced on the consumer side
            end = exc
        finally:
            # the decoders are stopped and JOINED before the end marker is published: the consumer (and with it the process)
            # may finish right after it, and a worker still inside cv2.VideoCapture.release() then aborts the interpreter
            self._stop_workers()
            if end is not None:
                self._q_end.append(end)
            self._finished.set()

    def _stop_workers(self):
        for ws in self._workers:
            for w in ws or ():
                w.tasks.put(None)
        for ws in self._workers:
            for w in ws or ():
                w.join()

    def close(self):
        """Stops and joins every background thread.  Called when the iteration ends for whatever reason (exhausted, `break`,
        exception); safe to call twice."""
        self._stop.set()
        for t in self._threads:
            if t is not threading.current_thread():
                t.join()
        self._stop_workers()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __iter__(self):
        try:
            while True:
                try:
                    item = self._q.get(timeout=0.05)
                except queue.Empty:
                    if self._finished.is_set() and self._q.empty():
                        if self._q_end:
                            raise self._q_end[0]
                        return
                    continue
                n, bufs = item
                yield n, [None if b is None else b[:n] for b in bufs]
                self._free.put(bufs)  # the consumer is done with the previous chunk when it asks for the next
        finally:
            self.close()


class DeviceChunkReader:
    """ChunkReader for clips this package wrote (ffv1_gpu.GpuFfv1Writer: FFV1 v3, every frame a key frame, ~1000 slices):
    the packets go from the file to the device and are decoded there, one thread per slice (`mdvt_ffv1_decode_frames`), so
    no decoded pixel ever exists on the host.  Same iteration protocol as ChunkReader -- (n, [tensor or None per path]) --
    but the tensors are CUDA tensors (n, H, W, 3) u8 RGB, or (n, H, W) grey for `grey` inputs (OpenCV's 8-bit BGR2GRAY,
    computed on the device).  A background thread stays `depth` chunks ahead on its own stream; a yielded chunk is
    complete and stays valid until the next one is asked for.  Raises `_lib.MdvtError` from `probe` / the constructor when
    a file is not such a stream (callers fall back to ChunkReader: open_chunk_reader)."""

    def __init__(self, paths: Sequence[Optional[str]], start: int = 0, stop: Optional[int] = None, chunk: int = 8, depth: int = 2,
                 grey: Sequence[bool] = (), device=None):
        from . import ffv1_gpu

        self.paths = list(paths)
        self.grey = list(grey) + [False] * (len(self.paths) - len(grey))
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.chunk = max(1, chunk)
        self._readers = [None if p is None else ffv1_gpu.GpuFfv1Reader(p, self.device, batch=self.chunk, rgb=True) for p in self.paths]
        first = next(r for r in self._readers if r is not None)
        self.width, self.height = first.width, first.height
        total = min(r.frames for r in self._readers if r is not None)
        self.start, self.stop = start, total if stop is None else min(stop, total)
        self._q: "queue.Queue" = queue.Queue(maxsize=max(1, depth))
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._run, name="mdvt-device-reader", daemon=True)
        self._thread.start()

    @staticmethod
    def _to_grey(rgb: torch.Tensor) -> torch.Tensor:
        """cv2.cvtColor(BGR2GRAY) of 8-bit frames: (R*9798 + G*19235 + B*3735 + 2^14) >> 15, OpenCV's 15-bit fixed-point
        weights (checked against cv2 4.13 on every third value of each channel: identical)."""
        x = rgb.to(torch.int32)
        return ((x[..., 0] * 9798 + x[..., 1] * 19235 + x[..., 2] * 3735 + 16384) >> 15).to(torch.uint8)

    def _run(self):
        from concurrent.futures import ThreadPoolExecutor

        streams = {}

        def decode_one(idx: int, a: int, b: int):
            """Frames [a, b) of input idx on that input's own stream: the inputs of a chunk are decoded side by side (the
            decoder is a latency-bound kernel with one thread per slice, two of them overlap almost perfectly)."""
            r, g = self._readers[idx], self.grey[idx]
            torch.cuda.set_device(self.device)
            if idx not in streams:
                streams[idx] = torch.cuda.Stream(device=self.device)
            with torch.cuda.stream(streams[idx]):
                frames = r.dec.decode([r._pk.payload(k) for k in range(a, b)], rgb=True)   # a fresh tensor, complete on return
                out = self._to_grey(frames) if g else frames
                streams[idx].synchronize()
            return out

        try:
            live = [k for k, r in enumerate(self._readers) if r is not None]
            with ThreadPoolExecutor(max_workers=max(1, len(live)), thread_name_prefix="mdvt-device-decode") as pool:
                for a in range(self.start, self.stop, self.chunk):
                    if self._stop.is_set():
                        return
                    b = min(a + self.chunk, self.stop)
                    futures = {k: pool.submit(decode_one, k, a, b) for k in live}
                    bufs = [futures[k].result() if k in futures else None for k in range(len(self._readers))]
                    while not self._stop.is_set():
                        try:
                            self._q.put((b - a, bufs), timeout=0.05)
                            break
                        except queue.Full:
                            continue
            self._q_put_end(None)
        except BaseException as exc:  # surfaced on the consumer side
            self._q_put_end(exc)

    def _q_put_end(self, item):
        while not self._stop.is_set():
            try:
                self._q.put(("end", item), timeout=0.05)
                return
            except queue.Full:
                continue

    def close(self):
        self._stop.set()
        if self._thread is not threading.current_thread():
            self._thread.join()
        for r in self._readers:
            if r is not None:
                r.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __iter__(self):
        try:
            while True:
                n, bufs = self._q.get()
                if n == "end":
                    if bufs is not None:
                        raise bufs
                    return
                yield n, bufs
        finally:
            self.close()


def open_chunk_reader(paths: Sequence[Optional[str]], start: int = 0, stop: Optional[int] = None, chunk: int = 8, depth: int = 3,
                      pin: bool = True, grey: Sequence[bool] = (), decoders: int = 1, device=None):
    """The reader of the script front ends: DeviceChunkReader when every input is an all-key-frame FFV1 stream of this
    package (the frames then never touch host memory), ChunkReader (cv2.VideoCapture threads, pinned host chunks) for
    everything else -- files written by the reference's tools included.  MDVT_FFV1_READER=host forces the latter."""
    if torch.cuda.is_available() and os.environ.get("MDVT_FFV1_READER", "") != "host":
        try:
            return DeviceChunkReader(paths, start, stop, chunk=chunk, depth=min(depth, 2), grey=grey, device=device)
        except Exception:  # noqa: BLE001 - not (all) device-decodable streams (another container, codec, slice layout, GOP > 1):
            pass           # the host reader takes over and reports whatever is really wrong with the files
    return ChunkReader(paths, start, stop, chunk=chunk, depth=depth, pin=pin, grey=grey, decoders=decoders)


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


GOP = 12  # key-frame interval of OpenCV's FFmpeg writer (AVCodecContext.gop_size): rank ranges start on key frames of the inputs


def gpu_ffv1_requested(flag: bool = False) -> bool:
    """FFV1 result videos coded on the device (ffv1_gpu.GpuFfv1Writer) instead of by cv2.VideoWriter lanes on host cores: the
    default wherever a CUDA device is present (the rendered frames then leave the device as packets only); a front end's
    --gpu_ffv1 or MDVT_FFV1_WRITER=gpu asks for it explicitly, MDVT_FFV1_WRITER=host keeps the host lanes."""
    env = os.environ.get("MDVT_FFV1_WRITER", "")
    if env == "host":
        return False
    return bool(flag) or env == "gpu" or torch.cuda.is_available()


def default_lanes(world_size: int = 1) -> int:
    env = os.environ.get("MDVT_WRITER_LANES")
    if env:
        return max(1, int(env))
    return max(1, min(12, (os.cpu_count() or 2) // (2 * max(1, world_size))))


class ParallelWriter:
    """An FFV1 .mkv written by up to `lanes` cv2.VideoWriters at once.  The single-threaded FFV1 entropy coder behind
    cv2.VideoWriter is the end-to-end limiter of every script (about 0.45 s per 3840x1080 frame); here the clip is cut
    into blocks of `block` frames, every block is encoded into its own small file by a worker thread (a fresh writer:
    the block starts with a key frame whatever its length), and `close()` stitches the encoded packets back into display
    order with `mkv_join.join` -- no re-encode, same container / codec parameters, frames decode bit-identically.
    Small blocks keep the drain short (a 96-frame clip is 16 blocks for 16 cores, not 8 GOPs); FFV1 is intra-only, a key
    frame merely resets the adaptive coder state, so the extra key frames cost about a per cent of file size.
    join_on_close=False leaves the block files and a `<path>.plan.json` for a later join of several writers' blocks
    (the torchrun ranks of one job)."""

    def __init__(self, path: str, fps: float, size: Tuple[int, int], lanes: Optional[int] = None, join_on_close: bool = True,
                 block: int = 4):
        from concurrent.futures import ThreadPoolExecutor

        self.path, self.fps, self.size, self.join_on_close = path, fps, size, join_on_close
        self.lanes = default_lanes() if lanes is None else max(1, lanes)
        self.block = max(1, block)
        self.plan: List[Tuple[str, int]] = []
        self.frames = 0
        self._pool = ThreadPoolExecutor(max_workers=self.lanes)
        self._slots = threading.Semaphore(2 * self.lanes)   # blocks encoded or waiting: bounds the host memory held
        self._futures = []
        self._pending: List[np.ndarray] = []
        self._pending_n = 0
        self._rgb: Optional[bool] = None

    def _encode(self, block_path: str, frames: np.ndarray, rgb: bool):
        cv2 = _cv2()
        try:
            w = cv2.VideoWriter(block_path, cv2.VideoWriter_fourcc(*"FFV1"), self.fps, self.size)
            if not w.isOpened():
                raise RuntimeError(f"cannot open an FFV1 writer for {block_path}")
            for f in frames:
                w.write(cv2.cvtColor(f, cv2.COLOR_RGB2BGR) if rgb else f)
            w.release()
        finally:
            self._slots.release()

    def _emit(self, n: int):
        """Hand the first n pending frames to a worker as one block."""
        take, got = [], 0
        while got < n:
            a = self._pending[0]
            need = n - got
            if a.shape[0] <= need:
                take.append(a)
                got += a.shape[0]
                self._pending.pop(0)
            else:
                take.append(a[:need])
                self._pending[0] = a[need:]
                got += need
        self._pending_n -= n
        frames = np.concatenate(take) if len(take) > 1 else take[0].copy()   # own copy: the caller recycles its buffer
        block_path = f"{self.path}.blk{len(self.plan):06d}.mkv"
        self.plan.append((block_path, n))
        self._slots.acquire()
        for f in self._futures:   # surface a worker's failure early
            if f.done() and f.exception() is not None:
                raise f.exception()
        self._futures.append(self._pool.submit(self._encode, block_path, frames, bool(self._rgb)))

    def write(self, frames, rgb: bool = True):
        arr = frames.numpy() if isinstance(frames, torch.Tensor) else np.asarray(frames)
        if arr.shape[1:3] != (self.size[1], self.size[0]):
            raise ValueError(f"frames are {arr.shape[2]}x{arr.shape[1]}, writer expects {self.size[0]}x{self.size[1]}")
        if self._rgb is None:
            self._rgb = rgb
        elif self._rgb != rgb:
            raise ValueError("one writer takes either RGB or BGR frames, not both")
        if arr.shape[0] == 0:
            return
        self._pending.append(arr)
        self._pending_n += arr.shape[0]
        self.frames += arr.shape[0]
        while self._pending_n >= self.block:
            self._emit(self.block)
        self._pending = [a.copy() for a in self._pending]   # the remainder outlives the caller's buffer

    def close(self):
        if self._pending_n:
            self._emit(self._pending_n)
        for f in self._futures:
            f.result()
        self._pool.shutdown()
        if self.join_on_close:
            join_plans([self.plan], self.path, self.fps)
        else:
            with open(self.path + ".plan.json", "w") as fh:
                json.dump({"fps": self.fps, "plan": self.plan}, fh)


def join_plans(plans: Sequence[Sequence[Tuple[str, int]]], out_path: str, fps: float, remove: bool = True) -> int:
    """Packet-level join of the lanes of one or more ParallelWriters (in the given order) into `out_path`."""
    from . import mkv_join

    plan = [(p, int(n)) for pl in plans for p, n in pl]
    lanes = sorted({p for p, _ in plan})
    total = mkv_join.join(plan, out_path, fps) if plan else 0
    if remove:
        for p in lanes:
            if os.path.exists(p):
                os.remove(p)
    return total


def load_plan(path: str, remove: bool = True):
    """The plan a ParallelWriter(join_on_close=False) left next to its block files."""
    with open(path + ".plan.json") as fh:
        d = json.load(fh)
    if remove:
        os.remove(path + ".plan.json")
    return [(p, int(n)) for p, n in d["plan"]]


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
