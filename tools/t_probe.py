"""Fused-kernel throughput per STFT frame as a function of the frame count T = 1 + n // 512 (8 warps per clip: T mod 8 frames
of the last round keep T mod 8 warps busy while the others wait at the barrier in front of the clip tail).
usage: python tools/t_probe.py [clips]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "multimodal-emotion-classification_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import bench  # noqa: E402
from sfx_b200 import get_extractor  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
dev = torch.device("cuda", 0)
ex = get_extractor(dev)
ex.set_pipeline("fused")
pool = bench.synth_pool(B, 66150 + 8 * 512, seed=7, device=dev)
for T in (120, 127, 128, 129, 130, 131, 132, 135, 136, 137):
    n = (T - 1) * 512 + 100
    w = pool[:, :n].contiguous()
    out = torch.empty((B, 56), device=dev)
    for _ in range(2):
        ex.extract(w, out=out)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(4):
        ex.extract(w, out=out)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 4
    print(f"T={T}: {B / ms / 1e3:.3f} M clips/s, {B * T / ms / 1e6:.1f} M frames/ms x1e-3, ns per frame-slot {ms * 1e6 / (B * T):.2f}", flush=True)
ex.set_pipeline("auto")
