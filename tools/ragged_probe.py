"""Throughput of a variable-length batch (lengths log-uniform in [0.5 s, 60 s]) against the
fixed-length rate: shows the cost of the longest clips landing late in the clip queue.
usage: python tools/ragged_probe.py [B] [max_seconds]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-emotion-classification_b200"))
from sfx_b200 import get_extractor

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
max_s = float(sys.argv[2]) if len(sys.argv) > 2 else 60.0
dev = torch.device("cuda", 0)
ex = get_extractor(dev)
rng = np.random.default_rng(0)
lens = np.exp(rng.uniform(np.log(11025), np.log(22050 * max_s), B)).astype(np.int32)
L = int(lens.max())
g = torch.Generator(device="cuda").manual_seed(1)
w = torch.randn((B, L), device=dev, generator=g) * 0.1
ld = torch.from_numpy(lens).to(dev)
frames = int((1 + lens // 512).sum())
out = torch.empty((B, 56), device=dev)


def run(order=None, reps=5):
    ww, ll = (w, ld) if order is None else (w[order], ld[order])
    for _ in range(2):
        ex.extract(ww, ll, out=out)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        ex.extract(ww, ll, out=out)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


for mode in (1, 3, 2):
    ex.lib.sfx_set_pipeline(mode)
    t = run()
    ts = run(torch.argsort(ld, descending=True))
    print(f"pipeline {('', 'fused', 'split', 'stream')[mode]}: B={B} frames={frames} as given {t*1e3:.2f} ms "
          f"({frames/t/130/1e6:.3f} M 3s-clip-equivalents/s), longest first {ts*1e3:.2f} ms ({frames/ts/130/1e6:.3f} M)")
ex.lib.sfx_set_pipeline(0)
