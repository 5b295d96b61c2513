"""Pinned host -> device copy bandwidth of this box (what bounds the end-to-end path): python tools/h2d_peak.py"""
import torch

dev = torch.device("cuda", 0)
for mb in (96, 256, 1024):
    h = torch.empty(mb * 1024 * 1024, dtype=torch.uint8).pin_memory()
    d = torch.empty_like(h, device=dev)
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    reps = 8
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    print(f"{mb} MiB x {reps}: {mb * 1024 * 1024 * reps / (a.elapsed_time(b) * 1e-3) / 1e9:.1f} GB/s", flush=True)
