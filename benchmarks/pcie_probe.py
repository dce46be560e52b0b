"""Host<->device copy ceilings of the box (pinned memory, large transfers): what bounds bench.py's `e2e` number.

    python benchmarks/pcie_probe.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 benchmarks/pcie_probe.py

Under torchrun every rank copies at the same time (barrier, CUDA-event timing, max over ranks), so the printed aggregate is
the ceiling the host side (PCIe switches, root complexes, host memory) gives N GPUs together -- the number the 1 -> 8 GPU
end-to-end scaling of bench.py has to be read against.  One JSON line on stdout (rank 0)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metric_depth_video_toolbox_b200 import sharding  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
locality = sharding.bind_host_to_gpu(local_rank, world) if "--no-bind" not in sys.argv else {"source": "not bound (--no-bind)"}
if world > 1:
    import torch.distributed as dist

    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)  # NCCL banners go to stderr
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

n = 256 << 20
host_a = torch.empty(n, dtype=torch.uint8, pin_memory=True)
host_b = torch.empty(n, dtype=torch.uint8, pin_memory=True)
dev_a = torch.empty(n, dtype=torch.uint8, device="cuda")
dev_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def sync_all():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def timed(fn, reps=8):
    fn()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    sync_all()
    t = torch.tensor([e0.elapsed_time(e1) / reps / 1e3], dtype=torch.float64, device="cuda")
    mine = float(t.item())
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return mine, float(t.item())


def both():
    with torch.cuda.stream(s1):
        dev_a.copy_(host_a, non_blocking=True)
    with torch.cuda.stream(s2):
        host_b.copy_(dev_b, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)


res = {}
for name, fn in (("h2d", lambda: dev_a.copy_(host_a, non_blocking=True)), ("d2h", lambda: host_b.copy_(dev_b, non_blocking=True)), ("both", both)):
    mine, slowest = timed(fn)
    res[name] = {"gbs_this_rank": n / mine / 1e9, "gbs_aggregate": world * n / slowest / 1e9}
# bench.py's e2e moves 12.44 MB in and 16.59 MB (u8 masks) / 12.96 MB (packed masks) out per 1080p stereo frame, both directions at once
agg = res["both"]["gbs_aggregate"]
res["e2e_ceiling_frames_per_s"] = {"u8_masks": agg * 1e9 / 16588800, "packed_masks": agg * 1e9 / 12960000,
                                   "how": "aggregate GB/s per direction with both directions busy / device-to-host bytes per frame (the larger direction)"}
per_rank = [None] * world
if world > 1:
    dist.all_gather_object(per_rank, {"rank": rank, "locality": locality, **{k: v["gbs_this_rank"] for k, v in res.items() if k != "e2e_ceiling_frames_per_s"}})
else:
    per_rank = [{"rank": 0, "locality": locality, **{k: v["gbs_this_rank"] for k, v in res.items() if k != "e2e_ceiling_frames_per_s"}}]
if rank == 0:
    line = {"probe": "pinned host <-> device copies, 256 MiB transfers, all ranks at once", "n_gpus": world,
            "aggregate": {k: (v["gbs_aggregate"] if isinstance(v, dict) and "gbs_aggregate" in v else v) for k, v in res.items()},
            "per_rank": per_rank}
    if world > 1:
        sys.stdout.flush()
        os.dup2(saved, 1)
    print(json.dumps(line), flush=True)
if world > 1:
    os.dup2(2, 1)
    dist.destroy_process_group()
