"""A/B of the tcgen05 chroma build (SFX_B200_LIB=ab/libumma.so): fused-kernel rows against the oracle on a small batch, then
throughput.  usage: SFX_B200_LIB=$PWD/ab/libumma.so python tools/umma_check.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "multimodal-emotion-classification_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import synth  # noqa: E402
from oracle import librosa_port as lp  # noqa: E402
from sfx_b200 import get_extractor  # noqa: E402

ex = get_extractor(torch.device("cuda", 0))
ex.set_pipeline("fused")
w = synth.make_batch(16, 66150, seed=3)
got = ex.extract(torch.from_numpy(w).cuda()).cpu().numpy()
torch.cuda.synchronize()
print("fused kernel returned", flush=True)
ok, rep = synth.compare(got, lp.features_batch(w))
print("parity vs oracle:", ok)
print(rep)
wr, lens = synth.make_ragged(12, 600, 200000, seed=3)
gr = ex.extract(torch.from_numpy(wr).cuda(), torch.from_numpy(lens).cuda()).cpu().numpy()
ok, rep = synth.compare(gr, lp.features_batch(wr, lens))
print("ragged parity:", ok)
print(rep)
