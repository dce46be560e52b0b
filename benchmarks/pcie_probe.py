"""Host<->device copy ceilings of the box (pinned memory, large transfers): what bounds bench.py's e2e number."""
import torch

n = 256 << 20
host_a = torch.empty(n, dtype=torch.uint8, pin_memory=True)
host_b = torch.empty(n, dtype=torch.uint8, pin_memory=True)
dev_a = torch.empty(n, dtype=torch.uint8, device="cuda")
dev_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=8):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps / 1e3


def both():
    with torch.cuda.stream(s1):
        dev_a.copy_(host_a, non_blocking=True)
    with torch.cuda.stream(s2):
        host_b.copy_(dev_b, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)


t = timed(lambda: dev_a.copy_(host_a, non_blocking=True))
print(f"H2D {n / t / 1e9:.1f} GB/s")
t = timed(lambda: host_b.copy_(dev_b, non_blocking=True))
print(f"D2H {n / t / 1e9:.1f} GB/s")
t = timed(both)
print(f"H2D+D2H concurrent {n / t / 1e9:.1f} GB/s each direction")
