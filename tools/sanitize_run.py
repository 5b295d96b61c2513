"""Tiny run for compute-sanitizer: a few clips of each kind (3 s + ragged lengths) through the extractor + DNN."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-emotion-classification_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth
from sfx_b200 import get_extractor

ex = get_extractor(torch.device("cuda", 0))
w = synth.make_batch(6, 66150, seed=1)
w[5, 40000:] = 0.0
out = ex.extract(torch.from_numpy(w).cuda())
wr, lens = synth.make_ragged(5, 600, 40000, seed=2)
out2 = ex.extract(torch.from_numpy(wr).cuda(), torch.from_numpy(lens).cuda())
torch.cuda.synchronize()
print("ok", float(out.sum()), float(out2.sum()))
