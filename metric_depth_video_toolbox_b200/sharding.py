"""Frame sharding across the GPUs of one box: one process per GPU, contiguous frame ranges,
and ONE small broadcast of the per-clip parameter block.  Pixels never cross GPUs.

The reference's analogue is scene-level subprocess parallelism (movie_2_3D.py:422-452).  Whole-clip
operations that span frames -- the Savitzky-Golay smoothing of the convergence list
(stereo_rerender.py:343-349), the lock-frame re-basing of the pose list (:369-373) -- are done on rank 0
*before* sharding and reach the other ranks only through `broadcast_params`.

Works with any torch.distributed backend: NCCL on the GPU box (tensors on the rank's device), gloo in the
CPU tests.  Without an initialised process group every function degrades to the single-rank case.
"""
from __future__ import annotations

import dataclasses
import json
import os
from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from .stereo import StereoParams


def world() -> Tuple[int, int]:
    """(rank, world_size) of the current process group, (0, 1) if there is none."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Join the torchrun job described by RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (no-op for a plain
    `python script.py`).  Returns (rank, world_size, local_rank)."""
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world_size > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    return rank, world_size, local_rank


def bind_host_to_gpu(local_rank: int, local_world: int = 1) -> dict:
    """Best-effort host locality for one-process-per-GPU jobs: the calling process is restricted to the CPU cores NVML
    reports as local to GPU `local_rank` (its PCIe root's socket), so that the pinned staging buffers it allocates
    afterwards are first touched -- and therefore placed -- on that socket's memory.  When NVML reports the same core
    set for every GPU (one NUMA node visible, as inside a VM) and several ranks share the box, the cores are split evenly
    between the ranks instead, so that the ranks' copy threads do not migrate over each other.  Returns what was done
    (reported by bench.py); never raises."""
    info = {"cores_before": None, "cores_after": None, "source": "unchanged"}
    try:
        before = sorted(os.sched_getaffinity(0))
        info["cores_before"] = len(before)
        local = None
        try:
            import pynvml

            pynvml.nvmlInit()
            handle = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
            words = pynvml.nvmlDeviceGetCpuAffinity(handle, (max(before) // 64) + 1)
            local = [c for c in before if (words[c // 64] >> (c % 64)) & 1]
            try:
                info["numa_node"] = int(pynvml.nvmlDeviceGetNumaNodeId(handle))
            except Exception:  # older drivers / VMs
                pass
        except Exception as exc:  # NVML not usable: keep the inherited mask
            info["nvml"] = f"{type(exc).__name__}"
        chosen, source = before, "unchanged"
        if local and len(local) < len(before):
            chosen, source = local, "nvml cpu affinity of the GPU"
        elif local_world > 1 and len(before) >= 2 * local_world:
            per = len(before) // local_world
            chosen, source = before[local_rank * per:(local_rank + 1) * per], "even split of the visible cores (no topology exposed)"
        if chosen != before:
            os.sched_setaffinity(0, set(chosen))
        info.update(cores_after=len(chosen), source=source)
    except Exception as exc:  # noqa: BLE001
        info["error"] = f"{type(exc).__name__}: {exc}"
    return info


def frame_range(n_frames: int, rank: Optional[int] = None, world_size: Optional[int] = None, align: int = 1) -> Tuple[int, int]:
    """Contiguous [start, stop) of `rank`: ranges differ by at most one frame (one block of `align` frames), concatenate
    in rank order to [0, n_frames), and are empty only when there are fewer frames (blocks) than ranks.  align > 1 puts
    every range start on a multiple of `align` (the key-frame interval of the clip files: seeks to a range start are
    exact and cheap, and every rank can decode with several threads)."""
    if rank is None or world_size is None:
        rank, world_size = world()
    if n_frames < 0 or world_size < 1 or not 0 <= rank < world_size or align < 1:
        raise ValueError(f"bad sharding request: {n_frames} frames, rank {rank} of {world_size}, align {align}")
    blocks = (n_frames + align - 1) // align
    base, extra = divmod(blocks, world_size)
    start = rank * base + min(rank, extra)
    stop = start + base + (1 if rank < extra else 0)
    return min(start * align, n_frames), min(stop * align, n_frames)


# ---------------------------------------------------------------------------------------------
# parameter block
# ---------------------------------------------------------------------------------------------
_SCALARS = ("width", "height", "xfov", "yfov", "max_depth", "pupillary_distance", "master_xfov", "near")
_FLAGS = ("infill_mask", "mask_rgb")
_NAN = float("nan")


def pack_params(p: StereoParams, n_frames: int) -> np.ndarray:
    """StereoParams -> one float64 vector: [n_frames, has_xfovs, has_conv, has_T, scalars..., flags...,
    xfovs (n), convergence (n), transforms (16 n)]; None scalars travel as NaN."""
    head = [float(n_frames), float(p.xfovs is not None), float(p.convergence_depths is not None), float(p.transformations is not None)]
    head += [_NAN if getattr(p, k) is None else float(getattr(p, k)) for k in _SCALARS]
    head += [float(bool(getattr(p, k))) for k in _FLAGS]
    parts = [np.asarray(head, dtype=np.float64)]
    for seq, per in ((p.xfovs, 1), (p.convergence_depths, 1), (p.transformations, 16)):
        if seq is not None:
            arr = np.asarray(seq, dtype=np.float64).reshape(-1)
            if arr.size != n_frames * per:
                raise ValueError(f"per-frame list has {arr.size // per} entries, clip has {n_frames} frames")
            parts.append(arr)
    return np.concatenate(parts)


def unpack_params(vec: np.ndarray) -> Tuple[StereoParams, int]:
    vec = np.asarray(vec, dtype=np.float64)
    n = int(vec[0])
    has_x, has_c, has_t = (bool(v) for v in vec[1:4])
    pos = 4
    kw = {}
    for k in _SCALARS:
        v = vec[pos]
        pos += 1
        kw[k] = None if np.isnan(v) else (int(v) if k in ("width", "height") else float(v))
    for k in _FLAGS:
        kw[k] = bool(vec[pos])
        pos += 1
    if has_x:
        kw["xfovs"] = vec[pos:pos + n].tolist()
        pos += n
    if has_c:
        kw["convergence_depths"] = vec[pos:pos + n].copy()
        pos += n
    if has_t:
        kw["transformations"] = vec[pos:pos + 16 * n].reshape(n, 4, 4).copy()
        pos += 16 * n
    return StereoParams(**kw), n


def broadcast_params(p, n_frames: int = 0, src: int = 0, device: Optional[torch.device] = None):
    """Rank `src` passes the clip's parameters; every rank returns (StereoParams, n_frames).  Two broadcasts:
    the vector length, then the vector (a few KB; at most ~140 B per frame with a pose file).
    `p` may also be the EXCEPTION rank `src` caught while building the parameters (a missing file, a bad list): the
    length travels as -1, every other rank raises too instead of waiting in the second broadcast until the NCCL
    timeout, and rank `src` re-raises the original."""
    rank, world_size = world()
    if world_size == 1:
        if isinstance(p, BaseException):
            raise p
        if p is None:
            raise ValueError("the source rank must supply the parameters")
        return p, n_frames
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    if rank == src:
        if p is None:
            p = ValueError("the source rank must supply the parameters")
        if isinstance(p, BaseException):
            size = torch.tensor([-1], dtype=torch.int64, device=device)
        else:
            vec = torch.from_numpy(pack_params(p, n_frames)).to(device)
            size = torch.tensor([vec.numel()], dtype=torch.int64, device=device)
    else:
        size = torch.zeros(1, dtype=torch.int64, device=device)
    dist.broadcast(size, src=src)
    if int(size.item()) < 0:
        if rank == src:
            raise p
        raise RuntimeError(f"rank {src} could not build the clip parameters (its exception carries the reason)")
    if rank != src:
        vec = torch.empty(int(size.item()), dtype=torch.float64, device=device)
    dist.broadcast(vec, src=src)
    return unpack_params(vec.cpu().numpy())


def shard_params(p: StereoParams, start: int, stop: int) -> StereoParams:
    """The parameters of frames [start, stop) only (per-frame lists sliced, everything else shared), for ranks
    that address their shard from frame 0."""
    cut = {}
    for k in ("xfovs", "convergence_depths", "transformations"):
        seq = getattr(p, k)
        cut[k] = None if seq is None else seq[start:stop]
    return dataclasses.replace(p, **cut)


def gather_counts(frames_done: int, device: Optional[torch.device] = None) -> int:
    """Sum of the per-rank frame counts (the closing all_reduce used for reporting)."""
    rank, world_size = world()
    if world_size == 1:
        return frames_done
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([frames_done], dtype=torch.int64, device=device)
    dist.all_reduce(t)
    return int(t.item())


def describe(n_frames: int, world_size: int) -> str:
    return json.dumps({"frames": n_frames, "ranks": world_size, "ranges": [frame_range(n_frames, r, world_size) for r in range(world_size)]})
