"""GPU probe of the stream pipeline (sfx_stream.cu): parity against the oracle, agreement with the fused kernel, and an A/B
of the three pipelines' device-resident throughput on the bench mix, per signal kind and as a whole.
usage: python tools/stream_probe.py [clips]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "multimodal-emotion-classification_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import bench  # noqa: E402
import synth  # noqa: E402
from oracle import librosa_port as lp  # noqa: E402
from sfx_b200 import get_extractor  # noqa: E402

dev = torch.device("cuda", 0)
ex = get_extractor(dev)
N = 66150


def cuda(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def run(mode, w, lens=None, debug=None):
    ex.set_pipeline(mode)
    try:
        out = ex.extract(w, lens, debug=debug)
        torch.cuda.synchronize()
        return out
    finally:
        ex.set_pipeline("auto")


# ---- 1. parity vs the oracle on small batches
w = synth.make_batch(24, N, seed=0)
ref = lp.features_batch(w)
dbg = {}
got = run("stream", cuda(w), debug=dbg).cpu().numpy()
ok, rep = synth.compare(got, ref)
print("stream vs oracle (24 clips, debug kernel):", ok)
print(rep)
print("fast-peak disagreements:", dbg["clip_info"].cpu().numpy()[:, 6].sum())
got2 = run("stream", cuda(w)).cpu().numpy()
print("debug == non-debug:", np.array_equal(got, got2))
edge = np.stack([synth.make_clip(k, N, np.random.default_rng(5)) for k in ("zero", "dc", "square")])
ok, rep = synth.compare(run("stream", cuda(edge)).cpu().numpy(), lp.features_batch(edge))
print("edge cases:", ok)
print(rep)
wr, lens = synth.make_ragged(12, 600, 200000, seed=3)
ok, rep = synth.compare(run("stream", cuda(wr), cuda(lens)).cpu().numpy(), lp.features_batch(wr, lens))
print("ragged:", ok)
print(rep)

# ---- 2. agreement with the fused kernel on a larger batch, determinism
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
pool = bench.synth_pool(B, N, seed=7, device=dev)
a = run("fused", pool[:4096])
b = run("stream", pool[:4096])
c = run("stream", pool[:4096])
d = (a - b).abs()
print("stream deterministic:", torch.equal(b, c))
print("stream vs fused: bitwise-equal columns:", [int(torch.equal(a[:, j], b[:, j])) for j in range(56)])
print("stream vs fused max abs diff per group: mfcc %.3g chroma %.3g zcr %.3g cent %.3g roll %.3g rms %.3g" % (
    d[:, :40].max(), d[:, 40:52].max(), d[:, 52].max(), d[:, 53].max(), d[:, 54].max(), d[:, 55].max()))


# ---- 3. throughput A/B
def rate(mode, w, reps):
    ex.set_pipeline(mode)
    out = torch.empty((w.shape[0], 56), device=dev)
    for _ in range(2):
        ex.extract(w, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        ex.extract(w, out=out)
    e1.record()
    torch.cuda.synchronize()
    ex.set_pipeline("auto")
    return w.shape[0] * reps / (e0.elapsed_time(e1) * 1e-3)


for rnd in range(2):
    for mode in ("fused", "stream"):
        kinds = [rate(mode, pool[k::4].contiguous(), 3) / 1e6 for k in range(4)]
        total = rate(mode, pool, 4) / 1e6
        print(f"round {rnd} {mode:6s}: noise {kinds[0]:.3f} harmonic {kinds[1]:.3f} noise_tail {kinds[2]:.3f} "
              f"harmonic_tail {kinds[3]:.3f} | mix {total:.3f} M clips/s", flush=True)
