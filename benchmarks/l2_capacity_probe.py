"""How much of a re-read buffer the B200's L2 keeps: read bandwidth of torch's sum over buffers of growing size, each read
50 times back to back (development probe for the z-buffer planes of the generic loop)."""
import torch
dev = torch.device("cuda:0")
for mb in (4, 8, 16, 24, 32, 40, 48, 64, 96, 128, 192, 512):
    x = torch.ones(mb * 1024 * 1024 // 8, dtype=torch.int64, device=dev)
    for _ in range(5):
        x.sum()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        x.sum()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    print(f"{mb:4d} MB re-read: {mb * 1.048576 / ms:8.1f} GB/s  ({ms * 1e3:.1f} us per pass)")
