"""Description of the polyphase resampler the device front-end runs (scope row f3).

It is scipy.signal.resample_poly(x, up, down) with its default Kaiser(5.0) low-pass, i.e. the resampler of this package's
``load_audio`` for files that are not at the target rate -- NOT librosa's soxr_hq, which cannot be restated or checked here.
The device kernel (csrc/sfx_frontend.cu) evaluates the same sums in the same order in float64, so its output is
bit-identical to ``resample_poly(x.astype(float64), up, down).astype(float32)``.
"""
from __future__ import annotations

import math

import numpy as np

_CACHE: dict = {}


def resample_filter(native_sr: int, target_sr: int) -> dict:
    """up, down, taps (float64, zero-padded as resample_poly pads them) and n_pre_remove for native_sr -> target_sr."""
    key = (int(native_sr), int(target_sr))
    if key in _CACHE:
        return _CACHE[key]
    g = math.gcd(key[1], key[0])
    up, down = key[1] // g, key[0] // g
    if up == down:
        out = dict(up=1, down=1, taps=np.zeros(1, dtype=np.float64), n_pre_remove=0)
    else:
        from scipy.signal import firwin          # scipy/signal/_signaltools.py::resample_poly, restated
        max_rate = max(up, down)
        half_len = 10 * max_rate
        h = firwin(2 * half_len + 1, 1.0 / max_rate, window=("kaiser", 5.0)) * up
        n_pre_pad = down - half_len % down
        out = dict(up=up, down=down, taps=np.ascontiguousarray(np.concatenate([np.zeros(n_pre_pad), h]), dtype=np.float64),
                   n_pre_remove=(half_len + n_pre_pad) // down)
    _CACHE[key] = out
    return out


def resample_poly_direct(x: np.ndarray, flt: dict) -> np.ndarray:
    """Plain-Python evaluation of the kernel's formula (float64 in, float64 out) -- test model for small inputs."""
    up, down, hp, npr = flt["up"], flt["down"], flt["taps"], flt["n_pre_remove"]
    x = np.asarray(x, dtype=np.float64)
    if up == down:
        return x.copy()
    n_in, L = len(x), len(hp)
    n_out = -(-n_in * up // down)
    y = np.zeros(n_out)
    for m in range(n_out):
        t0 = (m + npr) * down
        lo = max(0, -(-(t0 - (L - 1)) // up))
        hi = min(n_in - 1, t0 // up)
        acc = 0.0
        for i in range(lo, hi + 1):
            acc = acc + x[i] * hp[t0 - i * up]
        y[m] = acc
    return y
