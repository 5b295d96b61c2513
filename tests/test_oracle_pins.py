"""What can be pinned about the oracle without librosa itself (it is not installable here; tools/pin_against_librosa.py does
the rest wherever it is):

* librosa 0.10.0's two numba kernels on the path -- `_pi_stencil` / `_pi_wrapper` (core/pitch.py, the parabolic shift inside
  piptrack) and `_zc_stencil` / `_zc_wrapper` (core/audio.py, zero_crossings) -- written here with `numba.stencil` +
  `numba.guvectorize` as librosa writes them, so that numba's own type inference decides the float32 / float64 trail, against
  the oracle's vectorised restatements: bit for bit;
* scalar-loop re-derivations, straight from the formulas of SURVEY.md Appendix A (an independent form: one frame, one bin at a
  time, Python floats rounded where the spec says float32), of piptrack, estimate_tuning / pitch_tuning, spectral roll-off,
  rms and the zero-crossing rate, against the vectorised oracle.
"""
import math

import numpy as np
import pytest

import synth
from oracle import librosa_port as lp

SR = 22050
numba = pytest.importorskip("numba")


# ----------------------------------------------------------------------------------------------- numba kernels, as librosa types them
@numba.stencil
def _pi_stencil(x):
    a = x[1] + x[-1] - 2 * x[0]
    b = (x[1] - x[-1]) / 2
    if np.abs(b) >= np.abs(a):
        return 0
    return -b / a


@numba.guvectorize(["void(float32[:], float32[:])", "void(float64[:], float64[:])"], "(n)->(n)", cache=False, nopython=True)
def _pi_wrapper(x, y):
    y[:] = _pi_stencil(x)


def _parabolic_interpolation(x, axis=-2):
    xi = x.swapaxes(-1, axis)
    shifts = np.empty_like(x)
    shiftsi = shifts.swapaxes(-1, axis)
    _pi_wrapper(xi, shiftsi)
    shiftsi[..., -1] = 0
    shiftsi[..., 0] = 0
    return shifts


@numba.stencil
def _zc_stencil(x, threshold, zero_pos):
    x0 = x[0]
    if -threshold <= x0 <= threshold:
        x0 = 0
    x1 = x[-1]
    if -threshold <= x1 <= threshold:
        x1 = 0
    if zero_pos:
        return np.signbit(x0) != np.signbit(x1)
    else:
        return np.sign(x0) != np.sign(x1)


@numba.guvectorize(["void(float32[:], float32, bool_, bool_[:])", "void(float64[:], float64, bool_, bool_[:])"],
                   "(n),(),()->(n)", cache=False, nopython=True)
def _zc_wrapper(x, threshold, zero_pos, y):
    y[:] = _zc_stencil(x, threshold, zero_pos)


def _zero_crossings(y, threshold=1e-10, pad=True, zero_pos=True, axis=-1):
    yi = y.swapaxes(-1, axis)
    z = np.empty_like(y, dtype=bool)
    zi = z.swapaxes(-1, axis)
    _zc_wrapper(yi, threshold, zero_pos, zi)
    zi[..., 0] = pad
    return z


def _power_spectrogram(kind, n, seed):
    y = synth.make_clip(kind, n, np.random.default_rng(seed))
    return y, np.abs(lp.stft(y)) ** 2


def test_parabolic_shift_is_numbas_typing_of_the_librosa_stencil():
    """`2 * x[0]` and `/ 2` promote float32 to float64 inside the stencil, the sums stay float32, the store rounds back: the
    oracle's parabolic_shift states that trail by hand (and the CUDA kernel follows it); numba must agree bit for bit.
    An all-float32 reading is checked to differ, so the test can tell the two apart."""
    total = differs32 = 0
    for kind, seed in (("noise", 1), ("harmonic", 2), ("noise_tail", 3), ("square", 4)):
        _, P = _power_spectrogram(kind, 40000, seed)
        assert P.dtype == np.float32
        ref = _parabolic_interpolation(P, axis=-2)
        mine = lp.parabolic_shift(P)
        assert mine.dtype == np.float32 and np.array_equal(ref, mine), kind
        up, dn, mid = P[2:], P[:-2], P[1:-1]
        a32 = (up + dn) - np.float32(2) * mid
        b32 = (up - dn) / np.float32(2)
        with np.errstate(divide="ignore", invalid="ignore"):
            all32 = np.where(np.abs(b32) >= np.abs(a32), np.float32(0), -b32 / a32).astype(np.float32)
        differs32 += int((all32 != ref[1:-1]).sum())
        total += ref.size
        P64 = P.astype(np.float64)                                      # the float64 signature (float64 audio in the reference's tests)
        assert np.array_equal(_parabolic_interpolation(P64, axis=-2), lp.parabolic_shift(P64))
    assert total > 100_000 and differs32 > 1000


def test_zero_crossings_are_the_librosa_stencil():
    """zero_crossing_rate(y): edge pad, frames, zero_crossings(threshold=1e-10, zero_pos=True, pad=False) along the frame axis,
    mean.  numba types the threshold comparison of a float32 sample against float32(1e-10) for float32 audio."""
    rng = np.random.default_rng(5)
    for kind in ("noise", "harmonic", "dc", "square", "zero"):
        y = synth.make_clip(kind, 30000, rng)
        y[100:110] = 0.0
        y[200] = 5e-11
        y[201] = -5e-11
        y[202] = -2e-10
        ypad = np.pad(y, (1024, 1024), mode="edge")
        fr = lp.frame(ypad, 2048, 512)
        crossings = _zero_crossings(np.ascontiguousarray(fr), pad=False, axis=-2)
        ref = np.mean(crossings, axis=-2, keepdims=True)
        assert np.array_equal(ref, lp.zero_crossing_rate(y)), kind
        y64 = y.astype(np.float64)
        fr64 = lp.frame(np.pad(y64, (1024, 1024), mode="edge"), 2048, 512)
        ref64 = np.mean(_zero_crossings(np.ascontiguousarray(fr64), pad=False, axis=-2), axis=-2, keepdims=True)
        assert np.array_equal(ref64, lp.zero_crossing_rate(y64)), kind


# ----------------------------------------------------------------------------------------------- scalar re-derivations
def _scalar_piptrack(P):
    """App. A.3, one bin at a time: returns [(t, k, pitch float32, mag float32)] for every peak."""
    f32 = np.float32
    K, T = P.shape
    freq = [k * SR / 2048.0 for k in range(K)]
    peaks = []
    for t in range(T):
        col = [f32(v) for v in P[:, t]]
        ref = f32(f32(0.1) * max(col))
        masked = [v if v > ref else f32(0.0) for v in col]                 # S * (S > ref)
        for k in range(K):
            if not (150.0 <= freq[k] < 4000.0):
                continue
            left = masked[k - 1] if k > 0 else masked[0]                   # localmax pads with edge values
            right = masked[k + 1] if k < K - 1 else masked[K - 1]
            if not (masked[k] > left and masked[k] >= right):
                continue
            sm, sc, sp = col[k - 1], col[k], col[k + 1]                      # 14 <= k <= 371: interior
            a = float(f32(sp + sm)) - 2.0 * float(sc)                      # numba: float32 sum, float64 afterwards
            b = float(f32(sp - sm)) / 2.0
            shift = f32(0.0) if abs(b) >= abs(a) else f32(-b / a)
            avg = f32(f32(0.5) * f32(sp - sm))                             # np.gradient interior: (S[k+1] - S[k-1]) / 2
            dskew = f32(f32(f32(0.5) * avg) * shift)
            pitch = f32((k + float(shift)) * float(SR) / 2048)             # (int64 + float32 -> float64) * float / int
            mag = f32(sc + dskew)
            peaks.append((t, k, pitch, mag))
    return peaks


def _scalar_tuning(peaks):
    """App. A.3: median threshold over the peak magnitudes, then the 100-bin histogram arg-max of the residuals."""
    f32 = np.float32
    if not peaks:
        return 0.0, 0, 0.0
    mags = sorted(float(m) for _, _, _, m in peaks)
    n = len(mags)
    thr = f32(mags[n // 2]) if n % 2 else f32((f32(mags[n // 2 - 1]) + f32(mags[n // 2])) * f32(0.5))   # float32 mean of the middle pair
    edges = np.linspace(-0.5, 0.5, 101)
    counts = [0] * 100
    nsel = 0
    for _, _, pitch, mag in peaks:
        if not (mag >= thr and pitch > 0):
            continue
        nsel += 1
        octs = np.log2(f32(pitch) / f32(27.5))                            # float32 / python float (weak) -> float32 log2
        v = f32(12.0) * f32(octs)
        res = f32(np.mod(v, f32(1.0)))
        if res >= 0.5:
            res = f32(res - f32(1.0))
        r = float(res)
        b = None
        for i in range(100):                                              # np.histogram: [e_i, e_{i+1}), last bin closed
            if edges[i] <= r < edges[i + 1] or (i == 99 and r == edges[100]):
                b = i
                break
        if b is not None:
            counts[b] += 1
    best = max(range(100), key=lambda i: (counts[i], -i))                 # first arg-max
    return float(edges[best]), nsel, float(thr)


@pytest.mark.parametrize("kind,seed", [("noise", 11), ("harmonic", 12), ("harmonic_tail", 13)])
def test_piptrack_and_tuning_against_a_scalar_loop(kind, seed):
    _, P = _power_spectrogram(kind, 9000, seed)                           # 18 frames keep the Python loop short
    pitches, mags = lp.piptrack(P)
    peaks = _scalar_piptrack(P)
    assert len(peaks) == int((pitches > 0).sum()) > 0
    for t, k, pitch, mag in peaks:
        assert pitches[k, t] == pitch and mags[k, t] == mag, (t, k)
    tuning, dbg = lp.estimate_tuning(P, return_debug=True)
    t_ref, nsel, thr = _scalar_tuning(peaks)
    assert dbg["n_peaks"] == len(peaks) and dbg["n_sel"] == nsel and dbg["threshold"] == pytest.approx(thr, rel=0, abs=0)
    assert tuning == pytest.approx(t_ref, abs=1e-12)


def test_rolloff_rms_zcr_centroid_against_scalar_loops():
    f32 = np.float32
    rng = np.random.default_rng(21)
    for kind in ("noise", "harmonic", "noise_tail"):
        y = synth.make_clip(kind, 7000, rng)
        n, T = len(y), 1 + len(y) // 512
        S = np.abs(lp.stft(y))
        roll, cent = lp.spectral_rolloff(y)[0], lp.spectral_centroid(y)[0]
        rms, zcr = lp.rms(y)[0], lp.zero_crossing_rate(y)[0]
        assert len(roll) == len(cent) == len(rms) == len(zcr) == T
        for t in range(T):
            # roll-off: sequential float32 cumulative sum, first bin reaching 0.85 * total (all-zero frame: bin 0)
            c, cum = f32(0.0), []
            for k in range(1025):
                c = f32(c + S[k, t])
                cum.append(c)
            thr = float(f32(0.85) * cum[-1])                              # roll_percent * total_energy[-1]: a float32 array times a python float
            kk = next(k for k in range(1025) if not (float(cum[k]) < thr))
            assert roll[t] == kk * SR / 2048.0, (kind, t)
            # centroid: float64 frequencies times the magnitudes normalised by their float64 sum and stored as float32
            # (librosa.util.normalize returns the dtype of its input)
            tot = float(np.sum(S[:, t].astype(np.float64)))
            if tot < np.finfo(np.float32).tiny:
                tot = 1.0
            ref_c = math.fsum(k * SR / 2048.0 * float(f32(float(S[k, t]) / tot)) for k in range(1025))
            assert cent[t] == pytest.approx(ref_c, rel=1e-12, abs=1e-9)
            # rms: zero-padded frame, no window, float32 mean of squares
            seg = [float(y[i]) if 0 <= i < n else 0.0 for i in range(512 * t - 1024, 512 * t + 1024)]
            ref_r = math.sqrt(math.fsum(v * v for v in seg) / 2048.0)
            assert rms[t] == pytest.approx(ref_r, rel=2e-6, abs=1e-12)
            # zero-crossing rate: edge-padded frame, |x| <= 1e-10 counts as +0, position 0 of the frame never counts
            seg = [float(y[min(max(i, 0), n - 1)]) for i in range(512 * t - 1024, 512 * t + 1024)]
            sign = [(v < 0.0) and not (-1e-10 <= v <= 1e-10) for v in seg]
            ref_z = sum(1 for i in range(1, 2048) if sign[i] != sign[i - 1]) / 2048.0
            assert zcr[t] == ref_z, (kind, t)
