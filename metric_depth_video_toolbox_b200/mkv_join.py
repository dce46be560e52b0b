"""Packet-level concatenation of Matroska segments: the tail of the parallel FFV1 writers.

The reference writes every result video through ONE `cv2.VideoWriter(FFV1)` (stereo_rerender.py:420-442,941): about
0.45 s of single-threaded entropy coding per 3840x1080 side-by-side frame, i.e. ~2 frames/s end to end however fast the
per-pixel path is.  Here several `cv2.VideoWriter`s encode GOP-aligned blocks of the clip in parallel into lane files
(`video_io.ParallelWriter`, and one set of lanes per torchrun rank), and this module stitches the *encoded packets* of
the lanes back into one .mkv in display order without decoding anything: same container, same codec, same CodecPrivate
as a file written by a single writer, and the frames decode bit-identically (FFV1 is lossless; tests compare every
frame).  Only what the FFmpeg muxer behind OpenCV emits is understood: EBML header, one Segment with Info, Tracks (one
video track) and Clusters of SimpleBlocks / BlockGroups.
"""
from __future__ import annotations

import mmap
import os
import struct
from typing import BinaryIO, Iterator, List, Sequence, Tuple

ID_EBML, ID_SEGMENT = 0x1A45DFA3, 0x18538067
ID_INFO, ID_TRACKS, ID_CLUSTER, ID_CUES = 0x1549A966, 0x1654AE6B, 0x1F43B675, 0x1C53BB6B
ID_TIMECODE_SCALE, ID_DURATION, ID_MUXING_APP, ID_WRITING_APP = 0x2AD7B1, 0x4489, 0x4D80, 0x5741
ID_TIMECODE, ID_SIMPLE_BLOCK, ID_BLOCK_GROUP, ID_BLOCK, ID_REFERENCE_BLOCK = 0xE7, 0xA3, 0xA0, 0xA1, 0xFB
ID_CUE_POINT, ID_CUE_TIME, ID_CUE_TRACK_POSITIONS, ID_CUE_TRACK, ID_CUE_CLUSTER_POSITION = 0xBB, 0xB3, 0xB7, 0xF7, 0xF1
ID_TRACK_ENTRY, ID_CODEC_PRIVATE, ID_CODEC_ID, ID_DEFAULT_DURATION = 0xAE, 0x63A2, 0x86, 0x23E383
UNKNOWN_SIZE = -1


class MkvError(ValueError):
    pass


# ---- EBML primitives ---------------------------------------------------------------------------
def _read_id(buf: bytes, pos: int) -> Tuple[int, int]:
    first = buf[pos]
    length = 1
    mask = 0x80
    while length <= 4 and not first & mask:
        mask >>= 1
        length += 1
    if length > 4:
        raise MkvError(f"bad EBML id at {pos}")
    return int.from_bytes(buf[pos:pos + length], "big"), pos + length


def _read_size(buf: bytes, pos: int) -> Tuple[int, int]:
    first = buf[pos]
    length = 1
    mask = 0x80
    while length <= 8 and not first & mask:
        mask >>= 1
        length += 1
    if length > 8:
        raise MkvError(f"bad EBML size at {pos}")
    value = first & (mask - 1)
    for b in buf[pos + 1:pos + length]:
        value = (value << 8) | b
    if value == (1 << (7 * length)) - 1:
        value = UNKNOWN_SIZE
    return value, pos + length


def _enc_id(eid: int) -> bytes:
    return eid.to_bytes((eid.bit_length() + 7) // 8, "big")


def _enc_size(n: int, length: int = 0) -> bytes:
    if not length:
        length = 1
        while n >= (1 << (7 * length)) - 1:
            length += 1
    return ((1 << (7 * length)) | n).to_bytes(length, "big")


def _element(eid: int, payload: bytes) -> bytes:
    return _enc_id(eid) + _enc_size(len(payload)) + payload


def _uint(eid: int, v: int) -> bytes:
    return _element(eid, v.to_bytes(max(1, (v.bit_length() + 7) // 8), "big"))


def _children(buf: bytes, start: int, end: int) -> Iterator[Tuple[int, int, int]]:
    """(id, payload start, payload end) of the elements in buf[start:end]."""
    pos = start
    while pos < end:
        eid, p = _read_id(buf, pos)
        size, p = _read_size(buf, p)
        stop = end if size == UNKNOWN_SIZE else p + size
        if stop > end:
            raise MkvError(f"element 0x{eid:X} at {pos} runs past its parent")
        yield eid, p, stop
        pos = stop


# ---- reading one segment file -------------------------------------------------------------------
class MkvPackets:
    """The video packets of one OpenCV/FFmpeg-written .mkv, in file order: `packets` = [(offset, size, is_key)],
    plus the raw EBML header and Tracks element (codec parameters) for the output."""

    def __init__(self, path: str):
        self.path = path
        with open(path, "rb") as fh:   # mapped, not read: the lanes of a long clip add up to the size of the result
            buf = mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ) if os.path.getsize(path) else b""
        self._buf = buf
        top = list(_children(buf, 0, len(buf)))
        if not top or top[0][0] != ID_EBML:
            raise MkvError(f"{path}: not an EBML file")
        self.ebml_header = bytes(buf[0:top[0][2]])
        seg = next((t for t in top if t[0] == ID_SEGMENT), None)
        if seg is None:
            raise MkvError(f"{path}: no Segment")
        self.tracks = None
        self.timecode_scale = 1000000
        self.packets: List[Tuple[int, int, bool]] = []
        for eid, s, e in _children(buf, seg[1], seg[2]):
            if eid == ID_TRACKS:
                self.tracks = bytes(buf[s:e])
            elif eid == ID_INFO:
                for cid, cs, ce in _children(buf, s, e):
                    if cid == ID_TIMECODE_SCALE:
                        self.timecode_scale = int.from_bytes(buf[cs:ce], "big")
            elif eid == ID_CLUSTER:
                self._cluster(s, e)
        if self.tracks is None:
            raise MkvError(f"{path}: no Tracks element")

    def _block(self, s: int, e: int) -> Tuple[int, int, int]:
        """(payload offset, payload size, flags) of a (Simple)Block body: track vint, int16 timecode, flags, frame."""
        _, p = _read_size(self._buf, s)  # track number is coded like a size vint
        flags = self._buf[p + 2]
        if flags & 0x06:
            raise MkvError(f"{self.path}: laced blocks are not supported")
        return p + 3, e - (p + 3), flags

    def _cluster(self, start: int, end: int):
        for eid, s, e in _children(self._buf, start, end):
            if eid == ID_SIMPLE_BLOCK:
                off, size, flags = self._block(s, e)
                self.packets.append((off, size, bool(flags & 0x80)))
            elif eid == ID_BLOCK_GROUP:
                blk, key = None, True
                for cid, cs, ce in _children(self._buf, s, e):
                    if cid == ID_BLOCK:
                        blk = self._block(cs, ce)
                    elif cid == ID_REFERENCE_BLOCK:
                        key = False
                if blk is not None:
                    self.packets.append((blk[0], blk[1], key))

    def payload(self, index: int) -> bytes:
        off, size, _ = self.packets[index]
        return self._buf[off:off + size]

    def close(self):
        if isinstance(self._buf, mmap.mmap):
            self._buf.close()
        self._buf = b""

    def codec_private(self) -> bytes:
        for eid, s, e in _children(self.tracks, 0, len(self.tracks)):
            if eid == ID_TRACK_ENTRY:
                for cid, cs, ce in _children(self.tracks, s, e):
                    if cid == ID_CODEC_PRIVATE:
                        return self.tracks[cs:ce]
        return b""


# ---- writing the joined file --------------------------------------------------------------------
def _write_cluster(out: BinaryIO, t_ms: int, blocks: Sequence[Tuple[int, bool, bytes]]):
    body = [_uint(ID_TIMECODE, t_ms)]
    for rel, key, data in blocks:
        head = b"\x81" + struct.pack(">h", rel) + (b"\x80" if key else b"\x00")
        body.append(_enc_id(ID_SIMPLE_BLOCK) + _enc_size(len(head) + len(data)) + head)
        body.append(data)
    size = sum(len(b) for b in body)
    out.write(_enc_id(ID_CLUSTER) + _enc_size(size))
    for b in body:
        out.write(b)


def replace_codec_private(tracks: bytes, codec_private: bytes) -> bytes:
    """The Tracks payload with the (single) TrackEntry's CodecPrivate replaced; element sizes rebuilt."""
    out = b""
    for eid, s, e in _children(tracks, 0, len(tracks)):
        if eid != ID_TRACK_ENTRY:
            out += _enc_id(eid) + _enc_size(e - s) + tracks[s:e]
            continue
        entry = b""
        seen = False
        for cid, cs, ce in _children(tracks, s, e):
            if cid == ID_CODEC_PRIVATE:
                entry += _element(ID_CODEC_PRIVATE, codec_private)
                seen = True
            else:
                entry += _enc_id(cid) + _enc_size(ce - cs) + tracks[cs:ce]
        if not seen:
            entry += _element(ID_CODEC_PRIVATE, codec_private)
        out += _element(ID_TRACK_ENTRY, entry)
    return out


class StreamWriter:
    """One .mkv written packet by packet in display order: Info (Duration patched on close), the given Tracks payload,
    Clusters of SimpleBlocks that start on key frames, Cues per cluster.  The file appears under its name on close()."""

    def __init__(self, out_path: str, ebml_header: bytes, tracks: bytes, fps: float, frames_per_cluster: int = 24):
        self.out_path, self.tmp = out_path, out_path + ".joining"
        self.frame_ms = 1000.0 / fps
        self.frames_per_cluster = frames_per_cluster
        self.frames = 0
        self._cues: List[Tuple[int, int]] = []  # (time ms, cluster position relative to the Segment payload)
        self._pending: List[Tuple[int, bool, bytes]] = []
        self._cluster_t = 0
        self._out = open(self.tmp, "wb")
        out = self._out
        out.write(ebml_header)
        out.write(_enc_id(ID_SEGMENT) + _enc_size(0, 8))  # patched on close
        self._seg_start = out.tell()
        info = _uint(ID_TIMECODE_SCALE, 1000000) + _element(ID_MUXING_APP, b"mdvt-b200 mkv_join") + \
            _element(ID_WRITING_APP, b"mdvt-b200")
        duration = _element(ID_DURATION, struct.pack(">d", 0.0))
        out.write(_enc_id(ID_INFO) + _enc_size(len(info) + len(duration)) + info)
        self._duration_at = out.tell() + len(duration) - 8
        out.write(duration)
        out.write(_element(ID_TRACKS, tracks))

    def _flush(self):
        if self._pending:
            self._cues.append((self._cluster_t, self._out.tell() - self._seg_start))
            _write_cluster(self._out, self._cluster_t, self._pending)
            self._pending = []

    def add(self, data: bytes, key: bool = True):
        t = int(round(self.frames * self.frame_ms))
        if key and len(self._pending) >= self.frames_per_cluster or (self._pending and t - self._cluster_t > 30000):
            self._flush()
        if not self._pending:
            self._cluster_t = t
        self._pending.append((t - self._cluster_t, key, data))
        self.frames += 1

    def abort(self):
        if self._out is not None:
            self._out.close()
            self._out = None
            if os.path.exists(self.tmp):
                os.remove(self.tmp)

    def close(self) -> int:
        out = self._out
        self._flush()
        cue_body = b"".join(_element(ID_CUE_POINT, _uint(ID_CUE_TIME, t) + _element(
            ID_CUE_TRACK_POSITIONS, _uint(ID_CUE_TRACK, 1) + _uint(ID_CUE_CLUSTER_POSITION, pos))) for t, pos in self._cues)
        out.write(_element(ID_CUES, cue_body))
        seg_size = out.tell() - self._seg_start
        out.seek(self._seg_start - 8)
        out.write(_enc_size(seg_size, 8))
        out.seek(self._duration_at)
        out.write(struct.pack(">d", self.frames * self.frame_ms))
        out.close()
        self._out = None
        os.replace(self.tmp, self.out_path)
        return self.frames


def write_stream(out_path: str, ebml_header: bytes, tracks: bytes, packets, total: int, fps: float, frames_per_cluster: int = 24) -> int:
    """Write one .mkv from an iterable of (payload bytes, is_key) in display order (see StreamWriter)."""
    w = StreamWriter(out_path, ebml_header, tracks, fps, frames_per_cluster)
    try:
        for data, key in packets:
            w.add(data, key)
        if w.frames != total:
            raise MkvError(f"{out_path}: {w.frames} packets written, {total} expected")
        return w.close()
    except BaseException:
        w.abort()
        raise


def join(plan: Sequence[Tuple[str, int]], out_path: str, fps: float, frames_per_cluster: int = 24) -> int:
    """Write `out_path` from the packets of the lane files: `plan` = [(lane file, n packets), ...] in display order,
    each lane file consumed front to back across its entries.  Every run taken from a lane must start with a key
    frame (the lanes are filled in whole GOPs).  Returns the number of frames written."""
    lanes = {}
    cursor = {}
    for path, _ in plan:
        if path not in lanes:
            lanes[path] = MkvPackets(path)
            cursor[path] = 0
    first = lanes[plan[0][0]]
    for path, lane in lanes.items():
        if lane.tracks != first.tracks:
            if lane.codec_private() != first.codec_private():
                raise MkvError(f"{path}: codec parameters differ from {first.path}; the lanes cannot be joined at packet level")
    total = sum(n for _, n in plan)

    def packets():
        for path, n in plan:
            lane = lanes[path]
            at = cursor[path]
            if at + n > len(lane.packets):
                raise MkvError(f"{path}: {len(lane.packets)} packets, the plan asks for {at + n}")
            if n and not lane.packets[at][2]:
                raise MkvError(f"{path}: packet {at} starts a run but is not a key frame")
            for k in range(n):
                yield lane.payload(at + k), lane.packets[at + k][2]
            cursor[path] = at + n

    try:
        written = write_stream(out_path, first.ebml_header, first.tracks, packets(), total, fps, frames_per_cluster)
    finally:
        for lane in lanes.values():
            lane.close()
        if os.path.exists(out_path + ".joining"):
            os.remove(out_path + ".joining")
    return written
