"""Tiny run for compute-sanitizer (memcheck / racecheck / synccheck): a few clips of each kind (3 s, one with a zero tail,
plus a ragged batch) through every pipeline of the extractor, the host entry points (float32 and 16-bit PCM rows), the PCM
front-end (48 kHz stereo -> 22.05 kHz mono) and the scaler + DNN forward.

    compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_run.py [modes...] [--no-host] [--no-dnn]

modes default to: split fused fused_umma stream.  Prints one line per leg; rows of the pipelines are compared with the first
pipeline's (all pipelines share one arithmetic), so that a tool that perturbs scheduling and changes a row is seen.
(compute-sanitizer is closed on the B200 pool this was developed on -- it answers "closed on this pool" and exits 86 -- so the
committed evidence is the plain run, profiles/r02_all_entry_points_plain.txt; the command above is for a box that allows it.)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-emotion-classification_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import synth
from sfx_b200 import get_extractor

args = [a for a in sys.argv[1:] if not a.startswith("--")]
flags = {a for a in sys.argv[1:] if a.startswith("--")}
modes = args or ["split", "fused", "fused_umma", "stream"]

ex = get_extractor(torch.device("cuda", 0))
w = synth.make_batch(6, 66150, seed=1)
w[5, 40000:] = 0.0
wr, lens = synth.make_ragged(5, 600, 40000, seed=2)
wd, wrd, lend = torch.from_numpy(w).cuda(), torch.from_numpy(wr).cuda(), torch.from_numpy(lens).cuda()
first = None
for m in modes:
    ex.set_pipeline(m)
    out = ex.extract(wd).cpu().numpy()
    out2 = ex.extract(wrd, lend).cpu().numpy()
    torch.cuda.synchronize()
    assert np.isfinite(out).all() and np.isfinite(out2).all(), m
    if first is None:
        first = (out, out2)
    d = max(float(np.abs(out - first[0]).max()), float(np.abs(out2 - first[1]).max()))
    print(f"{m}: ok, fixed sum {float(out.sum()):.4f}, ragged sum {float(out2.sum()):.4f}, max |row - {modes[0]} row| {d:.3g}", flush=True)
ex.set_pipeline("auto")

if "--no-host" not in flags:
    h = ex.extract_host(w)
    print(f"extract_host f32: ok, max |row - device row| {float(np.abs(h - first[0]).max()) if modes[0] else 0:.3g}", flush=True)
    pcm = np.clip(np.round(w * 32767.0), -32768, 32767).astype(np.int16)
    hp = ex.extract_host(pcm)
    assert np.isfinite(hp).all()
    print(f"extract_host pcm16: ok, sum {float(hp.sum()):.4f}", flush=True)
    hr = ex.extract_host(wr, lens)
    print(f"extract_host ragged: ok, max |row - device row| {float(np.abs(hr - first[1]).max()):.3g}", flush=True)
    rng = np.random.default_rng(5)
    st = (rng.standard_normal((3, 2 * 48000 * 2)) * 6000).astype(np.int16)       # 2 s of 48 kHz stereo per clip
    fp = ex.preprocess_pcm16(st, np.array([96000, 70000, 12345]), 48000, channels=2)
    assert np.isfinite(fp).all()
    print(f"preprocess_pcm16 48 kHz stereo: ok, sum {float(fp.sum()):.4f}", flush=True)

if "--no-dnn" not in flags:
    from oracle import speech_dnn as od      # (weights generator only: tools/ is test infrastructure)
    from sfx_b200.dnn import SpeechDNN
    mdl = od.random_model(3)
    dnn = SpeechDNN(mdl, torch.device("cuda", 0))
    for B in (1, 7, 300):
        X = (mdl["scaler_mean"] + mdl["scaler_scale"] * np.random.default_rng(B).standard_normal((B, 56))).astype(np.float32)
        probs, tap = dnn.forward(torch.from_numpy(X).cuda())
        torch.cuda.synchronize()
        assert abs(float(probs.sum()) - B) < 1e-3 * B
    print("dnn forward: ok", flush=True)
print("done", flush=True)
