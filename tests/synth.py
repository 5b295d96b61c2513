"""Seeded synthetic waveforms and the parity comparator shared by tests, smoke() and bench.py.

Distributions (fp32 mono at 22 050 Hz, amplitudes in [-1, 1]):
  noise     white Gaussian, sigma = 0.1 (what the reference's own tests feed, tests/test_preprocessing.py:36)
  harmonic  speech-like harmonic stack: f0 in [90, 300] Hz with vibrato + glide, 20-25 harmonics ~1/h,
            AM envelope, noise floor at -50 dB
  tail      noise / harmonic with the last 20-45 % zeroed (the TESS-shaped case: top_db clamp active)
  zero / dc / square   edge cases
"""
import numpy as np

SR = 22050
KINDS = ("noise", "harmonic", "noise_tail", "harmonic_tail")

# |got - ref| <= RTOL * |ref| + ATOL[group]   (north_star: max relative error <= 1e-3 per pooled feature).
# The absolute terms cover features whose reference value is (near) zero, and are set from the float32
# self-noise of the reference arithmetic itself (DESIGN.md "Tolerance").
RTOL = 1e-3
ATOL = {"mfcc": 2e-3, "chroma": 1e-4, "zcr": 1e-7, "centroid": 1e-2, "rolloff": 1e-2, "rms": 1e-7}


def groups(n_mfcc=40):
    return {"mfcc": slice(0, n_mfcc), "chroma": slice(n_mfcc, n_mfcc + 12), "zcr": slice(n_mfcc + 12, n_mfcc + 13),
            "centroid": slice(n_mfcc + 13, n_mfcc + 14), "rolloff": slice(n_mfcc + 14, n_mfcc + 15),
            "rms": slice(n_mfcc + 15, n_mfcc + 16)}


def make_clip(kind, n, rng):
    t = np.arange(n, dtype=np.float64) / SR
    base = kind.split("_")[0]
    if base == "noise":
        y = 0.1 * rng.standard_normal(n)
    elif base == "harmonic":
        f0 = rng.uniform(90.0, 300.0)
        glide = rng.uniform(-0.15, 0.15) * f0
        vib = rng.uniform(0.0, 0.03) * f0
        fv = rng.uniform(4.0, 7.0)
        inst = f0 + glide * t / max(t[-1], 1e-9) + vib * np.sin(2 * np.pi * fv * t)
        phase = 2 * np.pi * np.cumsum(inst) / SR
        nh = int(rng.integers(20, 26))
        y = np.zeros(n)
        for h in range(1, nh + 1):
            if h * (f0 + abs(glide) + vib) < 0.5 * SR:
                y += np.sin(h * phase + rng.uniform(0, 2 * np.pi)) / h
        env = 0.55 + 0.45 * np.sin(2 * np.pi * rng.uniform(1.5, 4.0) * t + rng.uniform(0, 2 * np.pi))
        y = y * env
        y = 0.5 * y / max(np.abs(y).max(), 1e-9)
        y += 10 ** (-50 / 20) * 0.5 * rng.standard_normal(n)
    elif base == "zero":
        y = np.zeros(n)
    elif base == "dc":
        y = np.full(n, 0.25)
    elif base == "square":
        y = np.where(np.sin(2 * np.pi * 220.0 * t) >= 0, 1.0, -1.0)
    else:
        raise ValueError(kind)
    if kind.endswith("_tail"):
        cut = int(n * rng.uniform(0.55, 0.8))
        y[cut:] = 0.0
    return np.clip(y, -1.0, 1.0).astype(np.float32)


def make_batch(B, n, seed=0, kinds=KINDS):
    rng = np.random.default_rng(seed)
    out = np.empty((B, n), dtype=np.float32)
    for i in range(B):
        out[i] = make_clip(kinds[i % len(kinds)], n, rng)
    return out


def make_ragged(B, n_min, n_max, seed=0, kinds=KINDS):
    """Padded [B, n_max] batch with log-uniform lengths in [n_min, n_max] (config 5 shape)."""
    rng = np.random.default_rng(seed)
    lengths = np.exp(rng.uniform(np.log(n_min), np.log(n_max), size=B)).astype(np.int64)
    lengths[0], lengths[-1] = n_min, n_max
    out = np.zeros((B, n_max), dtype=np.float32)
    for i in range(B):
        out[i, :lengths[i]] = make_clip(kinds[i % len(kinds)], int(lengths[i]), rng)
        out[i, lengths[i]:] = 7.0          # poison the padding: it must never be read
    return out, lengths.astype(np.int32)


def compare(got, ref, n_mfcc=40, rtol=RTOL, atol=ATOL):
    """Returns (ok, text report).  Per group: max |err| / (rtol*|ref| + atol) must be <= 1."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    lines, ok = [], True
    for name, sl in groups(n_mfcc).items():
        g, r = got[:, sl], ref[:, sl]
        err = np.abs(g - r)
        budget = rtol * np.abs(r) + atol[name]
        ratio = err / budget
        rel = err / np.maximum(np.abs(r), 1e-30)
        worst = np.unravel_index(np.argmax(ratio), ratio.shape)
        fails = int((ratio > 1).any(axis=1).sum())
        ok &= fails == 0
        lines.append(f"{name:9s} max|err|={err.max():.3e} max err/budget={ratio.max():.3f} "
                     f"(clip {worst[0]}, got {g[worst]:.6g} ref {r[worst]:.6g}) "
                     f"median rel={np.median(rel):.2e} clips over budget={fails}/{len(g)}")
    return bool(ok), "\n".join(lines)
