"""Drop-in for the reference's `depth_frames_helper` module (depth_frames_helper.py:5-279): same function
names, argument order and results, with the per-pixel work done by libmdvt_b200's CUDA kernels.

Array arguments may be NumPy arrays (copied to the current CUDA device, result returned as NumPy, like the
reference) or CUDA tensors (zero-copy, result stays on the device).  There is no CPU implementation of the
codec here: without a CUDA device or without the library these functions raise.  The video-file helpers
(`save_depth_video`, `verify_and_move`, ...) keep OpenCV for container I/O exactly as the reference does and
only move the per-pixel encode onto the GPU.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import ops

# log-depth codec constants (depth_frames_helper.py:26-29)
C = 2.0
A = 16538.0


# ---------------------------------------------------------------------------------------------
# host <-> device plumbing
# ---------------------------------------------------------------------------------------------
def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("metric_depth_video_toolbox_b200 needs a CUDA device: there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def _up(a, dtype: torch.dtype):
    """(device tensor, came_from_numpy).  NumPy uint32 travels as its bit pattern."""
    if isinstance(a, torch.Tensor):
        if not a.is_cuda:
            a = a.to(_device())
        return a.to(dtype).contiguous(), False
    arr = np.ascontiguousarray(a)
    np_dtype = {torch.uint8: np.uint8, torch.float32: np.float32, torch.uint32: np.uint32, torch.float64: np.float64}[dtype]
    if arr.dtype != np_dtype:
        arr = arr.astype(np_dtype)
    if dtype == torch.uint32:
        return torch.from_numpy(arr.view(np.int32)).to(_device()).view(torch.uint32), True
    return torch.from_numpy(arr).to(_device()), True


def _down(t: torch.Tensor, as_numpy: bool):
    if not as_numpy:
        return t
    if t.dtype == torch.uint32:
        return t.view(torch.int32).cpu().numpy().view(np.uint32)
    return t.cpu().numpy()


# ---------------------------------------------------------------------------------------------
# linear codec (depth_frames_helper.py:5-24, 48-75, 99-103)
# ---------------------------------------------------------------------------------------------
def _is_f64(a) -> bool:
    return (a.dtype == torch.float64) if isinstance(a, torch.Tensor) else (np.asarray(a).dtype == np.float64)


def encode_depth_as_uint32(depth, max_depth):
    """clip to [0, max_depth] in the array's own dtype; (255**4 / max_depth * float64(depth)) truncated to uint32.
    float64 depth stays float64 (the reference never rounds it to float32); everything else is computed as float32."""
    d, as_np = _up(depth, torch.float64 if _is_f64(depth) else torch.float32)
    _, codes = ops.encode_depth(d, max_depth, True, True, want_codes=True)
    return _down(codes, as_np)


def decode_uint32_as_depth(encoded_value, max_depth):
    """float32(code) * float32(max_depth / 255**4) -> float32 metres."""
    codes, as_np = _up(encoded_value, torch.uint32)
    return _down(ops.codes_to_depth(codes, max_depth, "D1"), as_np)


def encode_data_as_BGR(data, frame_width, frame_height, bit16=False):
    """uint32 plane -> (H, W, 3) u8 in B, G, R order (16-bit: R = G = byte 3, B = byte 2)."""
    codes, as_np = _up(data, torch.uint32)
    codes = codes.reshape(frame_height, frame_width)
    return _down(ops.codes_to_pixels(codes, bit16, bgr_order=True), as_np)


def decode_rgb_as_data(rgb, frame_width, frame_height, bit16=False):
    """(H, W, 3) u8 RGB-order -> uint32 codes (16-bit: byte 3 <- R, byte 2 <- B; 24-bit: B, R, G into bytes 0-2)."""
    px, as_np = _up(rgb, torch.uint8)
    px = px.reshape(frame_height, frame_width, 3)
    return _down(ops.decode_depth(px, 100, bit16, "D1", want_codes=True, want_depth=False), as_np)


def decode_rgb_depth_frame(rgb, max_depth, bit16):
    """(H, W, 3) u8 RGB-order -> (H, W) float32 metres: one fused kernel."""
    px, as_np = _up(rgb, torch.uint8)
    return _down(ops.decode_depth(px, max_depth, bit16, "D1"), as_np)


# ---------------------------------------------------------------------------------------------
# log codec (depth_frames_helper.py:31-46) -- not on the per-frame path of any script; kept for API parity.
# Pointwise float64 transcendental work on host arrays; the scripts never call it.
# ---------------------------------------------------------------------------------------------
def encode_depth_as_uint32_log(depth, max_depth):
    depth = np.clip(depth, a_max=max_depth, a_min=0.0)
    return np.round(A * np.log1p(depth / C)).astype(np.uint32)


def decode_uint32_log_as_depth(encoded_value, max_depth):
    return (C * np.expm1(encoded_value.astype(np.float32) / A)).astype(np.float32)


# ---------------------------------------------------------------------------------------------
# image helpers (OpenCV, host side, as in the reference)
# ---------------------------------------------------------------------------------------------
def rescale_image(img, side_length, mode="max"):
    """depth_frames_helper.py:77-97."""
    import cv2

    h, w = img.shape[:2]
    if mode == "max":
        scale = side_length / max(h, w)
    elif mode == "min":
        scale = side_length / min(h, w)
    else:
        raise ValueError("mode must be 'max' or 'min'")
    return cv2.resize(img, (int(w * scale), int(h * scale)), interpolation=cv2.INTER_AREA)


def normalize_depth(d):
    """depth_frames_helper.py:105-123: 1st..99th percentile stretch to [0, 1]; None if nothing is finite."""
    d = d.astype(np.float32)
    finite = d[np.isfinite(d)]
    if finite.size == 0:
        return None
    lo, hi = np.percentile(finite, 1), np.percentile(finite, 99)
    if hi <= lo + 1e-6:
        return np.zeros_like(d, dtype=np.float32)
    return np.clip((d - lo) / (hi - lo), 0, 1).reshape(d.shape)


# ---------------------------------------------------------------------------------------------
# video files
# ---------------------------------------------------------------------------------------------
def _frame_geometry(frames):
    if isinstance(frames, np.ndarray):
        return frames.shape[0], frames.shape[1], frames.shape[2]
    return len(frames), frames[0].shape[0], frames[0].shape[1]


def save_depth_video(frames, output_video_path, fps, max_depth_arg, rescale_width, rescale_height):
    """depth_frames_helper.py:125-161: metric depth maps -> 16-bit RGB-coded FFV1 video.  Frames are encoded
    on the GPU in batches (clip, float64 scale, truncate, byte split in one kernel)."""
    import cv2

    nr_frames, height, width = _frame_geometry(frames)
    if isinstance(frames, np.ndarray):
        deepest = frames.max()
        print("max metric depth: ", deepest)
        if max_depth_arg < deepest:
            print("warning: output depth is deeper than max_depth. The depth will be clipped")
    from . import video_io

    # parallel FFV1 encoder lanes joined at packet level: the same frames as one cv2.VideoWriter would hold
    lanes = video_io.default_lanes()
    on_device = video_io.gpu_ffv1_requested()   # MDVT_FFV1_WRITER=gpu: wire-format pixels go from the encode kernel to the FFV1 coder
    if on_device:
        from . import ffv1_gpu

        out = ffv1_gpu.GpuFfv1Writer(output_video_path, fps, (rescale_width, rescale_height), device=_device())
    else:
        out = video_io.ParallelWriter(output_video_path, fps, (rescale_width, rescale_height), lanes=lanes) if lanes > 1 else \
            video_io.ChunkWriter(output_video_path, "FFV1", fps, (rescale_width, rescale_height))
    batch = max(1, min(nr_frames, (256 << 20) // max(1, rescale_width * rescale_height * 4)))
    for start in range(0, nr_frames, batch):
        chunk = []
        for i in range(start, min(nr_frames, start + batch)):
            depth = np.asarray(frames[i])
            if rescale_width != width or rescale_height != height:
                depth = cv2.resize(depth, (rescale_width, rescale_height), interpolation=cv2.INTER_LINEAR)
            chunk.append(np.ascontiguousarray(depth, dtype=np.float64 if depth.dtype == np.float64 else np.float32))
        dev = torch.from_numpy(np.stack(chunk)).to(_device())
        coded = ops.encode_depth(dev, max_depth_arg, True, True)  # B, G, R like encode_data_as_BGR
        out.write(coded if on_device else coded.cpu().numpy(), rgb=False)
    out.close()


def verify_and_move(tmp_file, expected_frames, output_file):
    """depth_frames_helper.py:163-179: promote a finished temporary video if its frame count is right."""
    import cv2

    if not os.path.isfile(tmp_file):
        return False
    cap = cv2.VideoCapture(tmp_file)
    if not cap.isOpened():
        return False
    actual_frames = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
    cap.release()
    if actual_frames != expected_frames:
        print(tmp_file, "not the correct nr of frames ", expected_frames, "!=", actual_frames)
        return False
    if os.path.exists(output_file):
        os.remove(output_file)
    os.rename(tmp_file, output_file)
    return True


def save_grayscale_video(frames, output_video_path, fps, max_depth_arg, rescale_width, rescale_height):
    """depth_frames_helper.py:181-232: depth -> 8-bit grey FFV1 preview (truncating cast, R = G = B)."""
    import cv2

    limit = float(max_depth_arg)
    nr_frames, height, width = _frame_geometry(frames)
    if isinstance(frames, np.ndarray) and limit < np.max(frames):
        print("warning: output depth exceeds max_depth_arg; values will be clipped.")
    out = cv2.VideoWriter(output_video_path, cv2.VideoWriter_fourcc(*"FFV1"), fps, (int(rescale_width), int(rescale_height)))
    for i in range(nr_frames):
        depth = frames[i]
        if depth.ndim == 3 and depth.shape[-1] == 1:
            depth = depth[..., 0]
        if rescale_width != width or rescale_height != height:
            depth = cv2.resize(depth, (int(rescale_width), int(rescale_height)), interpolation=cv2.INTER_LINEAR)
        denom = limit if limit > 0 else (depth.max() if np.max(depth) > 0 else 1.0)
        grey = ((np.clip(depth, 0, limit) / denom) * 255.0).astype(np.uint8)
        out.write(cv2.merge([grey, grey, grey]))
    out.release()


def write_video_frames_to_path(out_video, mask_frames, fps, H0, W0):
    """depth_frames_helper.py:234-249: RGB frames -> FFV1 file, nearest-neighbour resize when sizes differ."""
    import cv2

    writer = cv2.VideoWriter(out_video, cv2.VideoWriter_fourcc(*"FFV1"), fps, (W0, H0))
    assert writer.isOpened(), "Failed to open VideoWriter (FFV1/MKV). Try MJPG or mp4v if needed."
    for f in mask_frames:
        f = cv2.cvtColor(f, cv2.COLOR_RGB2BGR)
        if f.shape[0] != H0 or f.shape[1] != W0:
            f = cv2.resize(f, (W0, H0), interpolation=cv2.INTER_NEAREST)
        writer.write(f)
    writer.release()
    print(f"[ok] wrote {len(mask_frames)} frames to {out_video}")


def load_video_frames_from_path(video_path, start_frame=0, max_frames=-1):
    """depth_frames_helper.py:251-279: (list of RGB u8 frames, fps)."""
    import cv2

    if not os.path.exists(video_path):
        raise Exception("video file: " + video_path + " does not exist")
    cap = cv2.VideoCapture(video_path)
    assert cap.isOpened(), f"Failed to open video: {video_path}"
    fps = cap.get(cv2.CAP_PROP_FPS)
    frames = []
    idx = 0
    while True:
        ok, frame = cap.read()
        if not ok:
            break
        if idx >= start_frame:
            frames.append(cv2.cvtColor(frame, cv2.COLOR_BGR2RGB))
            if max_frames > 0 and len(frames) >= max_frames:
                break
        idx += 1
    cap.release()
    assert len(frames) > 0, "No frames read"
    return frames, fps
