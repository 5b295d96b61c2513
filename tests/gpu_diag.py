"""GPU triage script (not a test): device intermediates vs the oracle, plus a quick throughput probe.
Usage on the GPU box:  python tests/gpu_diag.py [B_timing]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa: F401
import synth
from oracle import librosa_port as lp
from sfx_b200 import get_extractor

torch.cuda.init()
print(torch.cuda.get_device_name(0))
ex = get_extractor(torch.device("cuda", 0))
print("grid_max workspace bytes (3 s):", ex.lib.sfx_workspace_bytes(0, 66150))

n = 66150
waves = synth.make_batch(8, n, seed=1)
waves = np.concatenate([waves, np.stack([synth.make_clip(k, n, np.random.default_rng(5)) for k in ("zero", "dc", "square")])])
dbg = {}
got = ex.extract(torch.from_numpy(waves).cuda(), debug=dbg).cpu().numpy()
torch.cuda.synchronize()
ref = lp.features_batch(waves)
ok, rep = synth.compare(got, ref)
print("features ok:", ok)
print(rep)
P = dbg["P"].cpu().numpy()[:, :, :1025]
LM = dbg["logmel"].cpu().numpy()
FF = dbg["frame_feat"].cpu().numpy()
CI = dbg["clip_info"].cpu().numpy()
for i in range(len(waves)):
    d = lp.debug_intermediates(waves[i])
    Pr = d["P"].T                                    # [T, 1025]
    fmax = np.maximum(Pr.max(axis=1, keepdims=True), 1e-30)
    eP = np.abs(P[i] - Pr) / fmax
    relP = np.abs(P[i] - Pr) / np.maximum(Pr, 1e-30)
    big = Pr > 1e-6 * fmax
    eL = np.abs(LM[i] - d["logmel"].T)
    cent = lp.spectral_centroid(waves[i])[0]
    roll = lp.spectral_rolloff(waves[i])[0]
    rms = lp.rms(waves[i])[0]
    print(f"clip {i}: P err/framemax {eP.max():.2e}, rel err on bins>1e-6max {relP[big].max() if big.any() else 0:.2e}; "
          f"logmel max abs err {eL.max():.2e}; cent {np.abs(FF[i,:,0]-cent).max():.2e} roll {np.abs(FF[i,:,1]-roll).max():.2e} "
          f"rms {np.abs(FF[i,:,2]-rms).max():.2e}; tuning gpu {CI[i,0]:+.2f} ref {d['tuning']:+.2f}; gmax {CI[i,1]:.4f}/{d['gmax']:.4f}; "
          f"npeaks {int(CI[i,2])}/{d['n_peaks']}; thr {CI[i,3]:.6g}/{d['threshold']:.6g}; nsel {int(CI[i,4])}/{d['n_sel']}")

# host path
got_h = ex.extract_host(waves)
print("host path == device path:", np.array_equal(got_h, got), np.abs(got_h - got).max())

# variable length
wr, lens = synth.make_ragged(12, 11025, 200000, seed=3)
got_r = ex.extract(torch.from_numpy(wr).cuda(), torch.from_numpy(lens).cuda()).cpu().numpy()
ref_r = lp.features_batch(wr, lens)
ok_r, rep_r = synth.compare(got_r, ref_r)
print("ragged ok:", ok_r)
print(rep_r)

# timing probe
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
g = torch.Generator(device="cuda").manual_seed(0)
big = torch.randn((B, n), device="cuda", generator=g) * 0.1
out = torch.empty((B, 56), device="cuda")
for _ in range(2):
    ex.extract(big, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    ex.extract(big, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"timing: B={B} {ms:.3f} ms/step -> {B / ms * 1e3:.0f} clips/s ({B / ms * 1e3 * 264824 / 1e9:.1f} GB/s algorithmic)")
