"""CPU oracle for the speech feature extractor -- TEST INFRASTRUCTURE ONLY.

*** PARITY UNPINNED ***  The arithmetic of the reference hot path
(/root/reference/preprocessing/audio_preprocessing.py:12-46) lives in the third-party
package ``librosa==0.10.0`` (reference requirements.txt:11; numpy==1.24.0 at :25) which is
not vendored under /root/reference and is not installable here (no wheel, no network).
The reference's own tests hold no golden vectors for this path (tests/test_preprocessing.py
:30-67 assert only shapes (40,), (12,), (4,) and finiteness).  This file therefore RESTATES the
published librosa 0.10.0 algorithms op for op (numpy + scipy), anchored on the reference's call
sites.  Independent cross-checks that do exist on this box are exercised in tests/test_oracle.py:
torchaudio's librosa-compatible Slaney mel bank / ortho DCT / MFCC, transformers.audio_utils'
"adapted from librosa" chroma bank and power_to_db, and hand-derived known answers (silence, DC,
bin-centred sinusoid).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module -- as the checker or the reported CPU baseline, never as the product
(the study scripts under ``tools/`` -- parity_study.py, pin_against_librosa.py, eq_check*.py -- are checkers of the same
kind: they ship no code path and nothing under the package imports them).

Each function cites the reference line it stands behind and the librosa 0.10.0 routine it restates.
dtype trail: float32 audio -> float64 window*frames -> float64 rfft -> complex64;
everything downstream float32 unless stated.  float64 audio keeps everything float64 (as librosa).
"""
from __future__ import annotations

import numpy as np
import scipy.fft
import scipy.fftpack

SR = 22050          # reference config.py:57
DURATION = 3        # reference config.py:58
N_MFCC = 40         # reference config.py:59
N_FFT = 2048        # librosa defaults
HOP = 512
N_MELS = 128
N_CHROMA = 12


class ParameterError(ValueError):
    """librosa.util.exceptions.ParameterError stand-in (reference lets it propagate)."""


# ----------------------------------------------------------------------------- util
def valid_audio(y):
    """librosa.util.valid_audio: ndarray, floating, finite."""
    if not isinstance(y, np.ndarray):
        raise ParameterError("Audio data must be of type numpy.ndarray")
    if not np.issubdtype(y.dtype, np.floating):
        raise ParameterError("Audio data must be floating-point")
    if y.ndim == 0 or y.shape[-1] == 0:
        raise ParameterError("Audio data must be at least one-dimensional and non-empty")
    if not np.isfinite(y).all():
        raise ParameterError("Audio buffer is not finite everywhere")
    return True


def tiny(x):
    """librosa.util.tiny: smallest normal of the (float) dtype of x."""
    x = np.asarray(x)
    if np.issubdtype(x.dtype, np.floating) or np.issubdtype(x.dtype, np.complexfloating):
        dtype = x.dtype
    else:
        dtype = np.dtype(np.float32)
    return np.finfo(dtype).tiny


def normalize(S, norm=np.inf, axis=0):
    """librosa.util.normalize (threshold=None, fill=None): lengths in float64, tiny -> 1."""
    threshold = tiny(S)
    mag = np.abs(S).astype(float)
    if norm == np.inf:
        length = np.max(mag, axis=axis, keepdims=True)
    elif norm == 1:
        length = np.sum(mag, axis=axis, keepdims=True)
    elif norm == 2:
        length = np.sum(mag ** 2, axis=axis, keepdims=True) ** 0.5
    else:
        raise ParameterError(f"unsupported norm {norm!r}")
    small_idx = length < threshold
    Snorm = np.empty_like(S)
    length[small_idx] = 1.0
    Snorm[:] = S / length
    return Snorm


def frame(y, frame_length=N_FFT, hop_length=HOP):
    """librosa.util.frame(axis=-1) on 1-D input: strided view [frame_length, n_frames]."""
    n = y.shape[-1]
    if n < frame_length:
        raise ParameterError(f"Input is too short (n={n}) for frame_length={frame_length}")
    n_frames = 1 + (n - frame_length) // hop_length
    s = y.strides[-1]
    return np.lib.stride_tricks.as_strided(
        y, shape=(frame_length, n_frames), strides=(s, s * hop_length), writeable=False)


def fft_frequencies(sr=SR, n_fft=N_FFT):
    """librosa.fft_frequencies = np.fft.rfftfreq(n_fft, 1/sr) (float64)."""
    return np.fft.rfftfreq(n=n_fft, d=1.0 / sr)


def hann_window(n_fft=N_FFT):
    """scipy.signal.get_window('hann', n_fft, fftbins=True): periodic Hann, float64."""
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n_fft) / n_fft)


# ----------------------------------------------------------------------------- STFT
def stft(y, n_fft=N_FFT, hop_length=HOP):
    """librosa.stft(center=True, pad_mode='constant', window='hann') -> [1+n_fft/2, T].

    float64 window * frames, float64 rfft, stored as complex64 for float32 input
    (librosa.util.dtype_r2c); complex128 for float64 input.  T = 1 + len(y)//hop.
    """
    valid_audio(y)
    win = hann_window(n_fft).reshape(-1, 1)
    ypad = np.pad(y, (n_fft // 2, n_fft // 2), mode="constant")
    frames = frame(ypad, n_fft, hop_length)
    cdtype = np.complex64 if y.dtype == np.float32 else np.complex128
    out = np.empty((1 + n_fft // 2, frames.shape[1]), dtype=cdtype, order="F")
    # librosa processes column blocks; the block size does not change the arithmetic
    blk = 256
    for s in range(0, frames.shape[1], blk):
        out[:, s:s + blk] = scipy.fft.rfft(win * frames[:, s:s + blk], axis=0)
    return out


def spectrogram(y, power, n_fft=N_FFT, hop_length=HOP):
    """librosa.core.spectrum._spectrogram: np.abs(stft)**power (power=1 -> abs only)."""
    S = np.abs(stft(y, n_fft=n_fft, hop_length=hop_length))
    if power != 1:
        S = S ** power
    return S


# ----------------------------------------------------------------------------- mel / MFCC
def hz_to_mel(f):
    """librosa.hz_to_mel(htk=False): Slaney scale."""
    f = np.asanyarray(f, dtype=float)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    if f.ndim:
        log_t = f >= min_log_hz
        mels[log_t] = min_log_mel + np.log(f[log_t] / min_log_hz) / logstep
    elif f >= min_log_hz:
        mels = min_log_mel + np.log(f / min_log_hz) / logstep
    return mels


def mel_to_hz(mels):
    """librosa.mel_to_hz(htk=False)."""
    mels = np.asanyarray(mels, dtype=float)
    f_sp = 200.0 / 3
    freqs = f_sp * mels
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    if mels.ndim:
        log_t = mels >= min_log_mel
        freqs[log_t] = min_log_hz * np.exp(logstep * (mels[log_t] - min_log_mel))
    elif mels >= min_log_mel:
        freqs = min_log_hz * np.exp(logstep * (mels - min_log_mel))
    return freqs


def mel_filterbank(sr=SR, n_fft=N_FFT, n_mels=N_MELS):
    """librosa.filters.mel(fmin=0, fmax=sr/2, htk=False, norm='slaney', dtype=float32)."""
    fmax = float(sr) / 2
    weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float32)
    fftfreqs = fft_frequencies(sr, n_fft)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(0.0), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]          # float32 *= float64 (computed in double, cast back)
    return weights


def power_to_db(S, ref=1.0, amin=1e-10, top_db=80.0):
    """librosa.power_to_db: 10 log10(max(amin,S)) - 10 log10(max(amin,ref)), clamp at max-top_db."""
    magnitude = np.asarray(S)
    log_spec = 10.0 * np.log10(np.maximum(amin, magnitude))
    log_spec -= 10.0 * np.log10(np.maximum(amin, np.abs(ref)))
    if top_db is not None:
        log_spec = np.maximum(log_spec, log_spec.max() - top_db)
    return log_spec


def melspectrogram(y, sr=SR):
    """librosa.feature.melspectrogram(power=2): einsum('...ft,mf->...mt', |stft|^2, mel_basis)."""
    S = spectrogram(y, power=2)
    return np.einsum("ft,mf->mt", S, mel_filterbank(sr), optimize=True)


def mfcc(y, sr=SR, n_mfcc=N_MFCC):
    """librosa.feature.mfcc(dct_type=2, norm='ortho', lifter=0) -> [n_mfcc, T]."""
    S = power_to_db(melspectrogram(y, sr))
    return scipy.fftpack.dct(S, axis=-2, type=2, norm="ortho")[:n_mfcc, :]


# ----------------------------------------------------------------------------- chroma / tuning
def hz_to_octs(frequencies, tuning=0.0, bins_per_octave=12):
    """librosa.hz_to_octs: log2(f / (A440/16)), A440 = 440*2^(tuning/bpo)."""
    A440 = 440.0 * 2.0 ** (tuning / bins_per_octave)
    return np.log2(np.asanyarray(frequencies) / (float(A440) / 16))


def chroma_filterbank(sr=SR, n_fft=N_FFT, tuning=0.0, n_chroma=N_CHROMA, ctroct=5.0, octwidth=2):
    """librosa.filters.chroma(norm=2, base_c=True, dtype=float32) -> [12, 1+n_fft/2]."""
    frequencies = np.linspace(0, sr, n_fft, endpoint=False)[1:]
    frqbins = n_chroma * hz_to_octs(frequencies, tuning=tuning, bins_per_octave=n_chroma)
    frqbins = np.concatenate(([frqbins[0] - 1.5 * n_chroma], frqbins))
    binwidthbins = np.concatenate((np.maximum(frqbins[1:] - frqbins[:-1], 1.0), [1]))
    D = np.subtract.outer(frqbins, np.arange(0, n_chroma, dtype="d")).T
    n_chroma2 = np.round(float(n_chroma) / 2)
    D = np.remainder(D + n_chroma2 + 10 * n_chroma, n_chroma) - n_chroma2
    wts = np.exp(-0.5 * (2 * D / np.tile(binwidthbins, (n_chroma, 1))) ** 2)
    wts = normalize(wts, norm=2, axis=0)
    wts *= np.tile(np.exp(-0.5 * (((frqbins / n_chroma - ctroct) / octwidth) ** 2)), (n_chroma, 1))
    wts = np.roll(wts, -3 * (n_chroma // 12), axis=0)
    return np.ascontiguousarray(wts[:, :int(1 + n_fft / 2)], dtype=np.float32)


def parabolic_shift(S):
    """librosa.core.pitch._parabolic_interpolation (numba stencil) along axis 0.

    numba typing: ``2 * x[0]`` and ``(...) / 2`` promote float32 to float64, while
    ``x[1] + x[-1]`` and ``x[1] - x[-1]`` stay in the input dtype; the result is cast back
    to the input dtype on store.  Edges are 0.
    """
    S = np.asarray(S)
    up, dn, mid = S[2:], S[:-2], S[1:-1]
    a = (up + dn).astype(np.float64) - 2.0 * mid.astype(np.float64)
    b = (up - dn).astype(np.float64) / 2.0
    with np.errstate(divide="ignore", invalid="ignore"):
        sh = np.where(np.abs(b) >= np.abs(a), 0.0, -b / a)
    out = np.zeros_like(S)
    out[1:-1] = sh.astype(S.dtype)
    return out


def localmax(x):
    """librosa.util.localmax(axis=0): x > x[k-1] (edge-padded) and x >= x[k+1]."""
    xp = np.pad(x, [(1, 1)] + [(0, 0)] * (x.ndim - 1), mode="edge")
    return (x > xp[:-2]) & (x >= xp[2:])


def piptrack(S, sr=SR, n_fft=N_FFT, fmin=150.0, fmax=4000.0, threshold=0.1):
    """librosa.piptrack(S=S) -- S is used as given (chroma_stft hands it the POWER spectrogram)."""
    S = np.abs(S)
    fmin = np.maximum(fmin, 0)
    fmax = np.minimum(fmax, float(sr) / 2)
    fft_freqs = fft_frequencies(sr, n_fft)
    avg = np.gradient(S, axis=-2)
    shift = parabolic_shift(S)
    dskew = 0.5 * avg * shift
    pitches = np.zeros_like(S)
    mags = np.zeros_like(S)
    freq_mask = ((fmin <= fft_freqs) & (fft_freqs < fmax)).reshape(-1, 1)
    ref_value = threshold * np.max(S, axis=-2)
    ref_value = np.expand_dims(ref_value, -2)
    idx = np.nonzero(freq_mask & localmax(S * (S > ref_value)))
    pitches[idx] = (idx[-2] + shift[idx]) * float(sr) / n_fft
    mags[idx] = S[idx] + dskew[idx]
    return pitches, mags


def pitch_tuning(frequencies, resolution=0.01, bins_per_octave=12):
    """librosa.pitch_tuning: histogram arg-max of the fractional-semitone residuals."""
    frequencies = np.atleast_1d(frequencies)
    frequencies = frequencies[frequencies > 0]
    if not np.any(frequencies):
        return 0.0          # librosa warns "Trying to estimate tuning from empty frequency set."
    residual = np.mod(bins_per_octave * hz_to_octs(frequencies, bins_per_octave=bins_per_octave), 1.0)
    residual[residual >= 0.5] -= 1.0
    bins = np.linspace(-0.5, 0.5, int(np.ceil(1.0 / resolution)) + 1)
    counts, tuning = np.histogram(residual, bins)
    return float(tuning[np.argmax(counts)])


def estimate_tuning(S, sr=SR, n_fft=N_FFT, resolution=0.01, bins_per_octave=12, return_debug=False):
    """librosa.estimate_tuning(S=S): piptrack -> median magnitude threshold -> pitch_tuning."""
    pitch, mag = piptrack(S, sr=sr, n_fft=n_fft)
    pitch_mask = pitch > 0
    if pitch_mask.any():
        threshold = np.median(mag[pitch_mask])
    else:
        threshold = 0.0
    sel = pitch[(mag >= threshold) & pitch_mask]
    t = pitch_tuning(sel, resolution=resolution, bins_per_octave=bins_per_octave)
    if return_debug:
        return t, dict(n_peaks=int(pitch_mask.sum()), threshold=float(threshold), n_sel=int(sel.size))
    return t


def chroma_stft(y, sr=SR, tuning=None):
    """librosa.feature.chroma_stft(norm=inf, n_chroma=12, tuning=None) -> [12, T]."""
    S = spectrogram(y, power=2)
    if tuning is None:
        tuning = estimate_tuning(S, sr=sr, bins_per_octave=N_CHROMA)
    chromafb = chroma_filterbank(sr, N_FFT, tuning=tuning)
    raw = np.einsum("cf,ft->ct", chromafb, S, optimize=True)
    return normalize(raw, norm=np.inf, axis=-2)


# ----------------------------------------------------------------------------- spectral
def zero_crossing_rate(y, frame_length=N_FFT, hop_length=HOP):
    """librosa.feature.zero_crossing_rate: EDGE pad, zero_crossings(threshold=1e-10, zero_pos, pad=False)."""
    valid_audio(y)
    ypad = np.pad(y, (frame_length // 2, frame_length // 2), mode="edge")
    fr = frame(ypad, frame_length, hop_length)
    x = np.where(np.abs(fr) <= 1e-10, 0, fr)       # -thr <= x <= thr -> 0 (positive zero)
    sb = np.signbit(x)
    z = np.empty(fr.shape, dtype=bool)
    z[1:] = sb[1:] != sb[:-1]
    z[0] = False                                   # pad=False
    return np.mean(z, axis=-2, keepdims=True)


def spectral_centroid(y, sr=SR):
    """librosa.feature.spectral_centroid: sum(freq64 * normalize(|stft|, norm=1))."""
    S = spectrogram(y, power=1)
    freq = fft_frequencies(sr).reshape(-1, 1)
    return np.sum(freq * normalize(S, norm=1, axis=-2), axis=-2, keepdims=True)


def spectral_rolloff(y, sr=SR, roll_percent=0.85):
    """librosa.feature.spectral_rolloff: first bin where cumsum(|stft|) >= 0.85 * total."""
    S = spectrogram(y, power=1)
    freq = fft_frequencies(sr).reshape(-1, 1)
    total_energy = np.cumsum(S, axis=-2)
    threshold = roll_percent * total_energy[-1, :]
    threshold = np.expand_dims(threshold, -2)
    ind = np.where(total_energy < threshold, np.nan, 1)
    return np.nanmin(ind * freq, axis=-2, keepdims=True)


def rms(y, frame_length=N_FFT, hop_length=HOP):
    """librosa.feature.rms(center=True, pad_mode='constant'): sqrt(mean(y^2)) per frame, no window."""
    valid_audio(y)
    ypad = np.pad(y, (frame_length // 2, frame_length // 2), mode="constant")
    x = frame(ypad, frame_length, hop_length)
    power = np.mean(np.abs(x) ** 2, axis=-2, keepdims=True)
    return np.sqrt(power)


# ----------------------------------------------------------------------------- the reference module's functions
def pad_or_trim(audio, sr=SR, duration=DURATION):
    """reference audio_preprocessing.py:14-18 (the part of load_audio after librosa.load)."""
    target_len = sr * duration
    if len(audio) < target_len:
        audio = np.pad(audio, (0, target_len - len(audio)), mode="constant")
    else:
        audio = audio[:target_len]
    return audio


def extract_mfcc(audio, sr=SR, n_mfcc=N_MFCC):
    """reference audio_preprocessing.py:22-24."""
    m = mfcc(audio, sr=sr, n_mfcc=n_mfcc)
    return np.mean(m.T, axis=0)


def extract_chroma(audio, sr=SR):
    """reference audio_preprocessing.py:27-29."""
    c = chroma_stft(audio, sr=sr)
    return np.mean(c.T, axis=0)


def extract_spectral_features(audio, sr=SR):
    """reference audio_preprocessing.py:32-37."""
    zcr = float(np.mean(zero_crossing_rate(audio)))
    centroid = float(np.mean(spectral_centroid(audio, sr=sr)))
    rolloff = float(np.mean(spectral_rolloff(audio, sr=sr)))
    r = float(np.mean(rms(audio)))
    return np.array([zcr, centroid, rolloff, r], dtype=np.float32)


def features_from_audio(audio, sr=SR):
    """reference audio_preprocessing.py:42-46 (preprocess_audio minus the file decode)."""
    feats = np.concatenate([extract_mfcc(audio, sr), extract_chroma(audio, sr),
                            extract_spectral_features(audio, sr)])
    return feats.astype(np.float32)


def features_batch(waves, lengths=None, sr=SR):
    """Loop of features_from_audio over rows (row i uses waves[i, :lengths[i]])."""
    out = np.empty((len(waves), 56), dtype=np.float32)
    for i, w in enumerate(waves):
        n = len(w) if lengths is None else int(lengths[i])
        out[i] = features_from_audio(np.ascontiguousarray(w[:n]), sr)
    return out


def debug_intermediates(audio, sr=SR):
    """Per-clip intermediates for parity triage (one shared STFT; same arithmetic)."""
    X = stft(audio)
    P = np.abs(X) ** 2
    S = np.abs(X)
    mel = np.einsum("ft,mf->mt", P, mel_filterbank(sr), optimize=True)
    tuning, dbg = estimate_tuning(P, sr=sr, bins_per_octave=N_CHROMA, return_debug=True)
    logmel = 10.0 * np.log10(np.maximum(1e-10, mel))
    return dict(P=P, S=S, mel=mel, logmel=logmel, gmax=float(logmel.max()), tuning=tuning, **dbg)
