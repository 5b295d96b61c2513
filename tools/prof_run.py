"""Short single-GPU run for ncu: a few launches of the extractor kernel on B synthetic 3 s clips."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-emotion-classification_b200"))
from sfx_b200 import get_extractor

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ex = get_extractor(torch.device("cuda", 0))
g = torch.Generator(device="cuda").manual_seed(0)
w = torch.randn((B, 66150), device="cuda", generator=g) * 0.1
out = torch.empty((B, 56), device="cuda")
for _ in range(iters):
    ex.extract(w, out=out)
torch.cuda.synchronize()
print("ok", float(out.sum()))
