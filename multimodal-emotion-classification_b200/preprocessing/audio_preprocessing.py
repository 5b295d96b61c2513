"""
Audio preprocessing utilities -- B200-native drop-in for the reference module of the same name
(/root/reference/preprocessing/audio_preprocessing.py:12-46).

Same five entry points, argument names, defaults and return shapes: fixed-duration loading (zero pad / trim) and the
frame-mean MFCC, chroma and [zcr, centroid, rolloff, rms] descriptors.

The arithmetic that the reference delegates to librosa (STFT, mel/log/DCT, tuning estimate + chroma,
zcr/centroid/rolloff/rms, frame-mean pooling) runs in hand-written sm_100a CUDA kernels behind the C ABI of
include/sfx.h (libsfx_b200.so).  There is no CPU fallback: importing this module needs no GPU, calling an
extract_* function without one raises ``sfx_b200.NoCudaDeviceError``.

Additive batched entry points: ``extract_features_batch`` and ``preprocess_audio_batch``.
"""

import os
import struct
import sys
import zlib

import numpy as np

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG_ROOT not in sys.path:
    sys.path.insert(0, _PKG_ROOT)

from sfx_b200._config import Config  # noqa: E402  (reference :9: `from config import Config`; the reference's wins when importable)


class ParameterError(ValueError):
    """Raised for invalid audio buffers, like librosa.util.exceptions.ParameterError in the reference."""


# --------------------------------------------------------------------------- helpers
def _valid_audio(audio):
    """librosa.util.valid_audio semantics: ndarray, floating point, non-empty, finite everywhere."""
    if not isinstance(audio, np.ndarray):
        raise ParameterError("Audio data must be of type numpy.ndarray")
    if not np.issubdtype(audio.dtype, np.floating):
        raise ParameterError("Audio data must be floating-point")
    if audio.ndim != 1 or audio.shape[0] == 0:
        raise ParameterError("Audio data must be a non-empty 1-D array (mono)")
    if not np.isfinite(audio).all():
        raise ParameterError("Audio buffer is not finite everywhere")


_last = (None, None)       # (key, row) of the most recent clip; replaced in one assignment, so threads never see a mixed pair


def _features_1clip(audio, sr, n_mfcc):
    """Run the fused kernel on one clip; the three extract_* calls the reference makes back to back on the
    same buffer (audio_preprocessing.py:42-44) share one launch through a content-keyed single-entry cache."""
    _valid_audio(audio)
    a32 = np.ascontiguousarray(audio, dtype=np.float32)
    key = (a32.shape[0], int(sr), int(n_mfcc), zlib.crc32(a32.view(np.uint8)))
    global _last
    last_key, row = _last
    if last_key != key:
        from sfx_b200 import get_extractor
        ex = get_extractor(None, int(sr))
        row = ex.extract_host(a32.reshape(1, -1), n_mfcc=int(n_mfcc))[0]
        _last = (key, row)
    return row, np.result_type(audio.dtype, np.float32)


class UnsupportedAudioFormat(ValueError):
    """The file is not something the built-in RIFF/WAVE reader decodes (mp3 / ogg / flac, compressed WAVE tags) and no host
    decoder is importable.  The reference's librosa.load goes through soundfile / audioread for these (config.py:49 lets
    the Flask app accept mp3 and ogg); this drop-in uses the same two packages when they are installed and otherwise
    raises this ValueError subclass, which the reference's routes report like any other unreadable upload."""


def _host_decode(file_path, why):
    """(float32 [frames, channels], rate) through soundfile, else audioread -- the decoders behind librosa.load."""
    try:
        import soundfile
        x, rate = soundfile.read(file_path, dtype="float32", always_2d=True)
        return np.ascontiguousarray(x), int(rate)
    except ImportError:
        pass
    try:
        import audioread
        with audioread.audio_open(file_path) as fh:
            rate, channels = int(fh.samplerate), int(fh.channels)
            pcm = np.frombuffer(b"".join(fh), dtype="<i2")
        return (pcm[:len(pcm) // channels * channels].reshape(-1, channels).astype(np.float32) / 32768.0), rate
    except ImportError:
        raise UnsupportedAudioFormat(f"{file_path}: {why}; install soundfile or audioread to decode it on the host") from None


def _wav_chunks(file_path):
    """(fmt tuple, raw data bytes) of a RIFF/WAVE file."""
    with open(file_path, "rb") as fh:
        data = fh.read()
    if len(data) < 12 or data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise UnsupportedAudioFormat(f"{file_path}: not a RIFF/WAVE file (only WAV decoding is built in)")
    pos, fmt, raw = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack("<I", data[pos + 4:pos + 8])[0]
        body = data[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            fmt = struct.unpack("<HHIIHH", body[:16])
            if fmt[0] == 0xFFFE and len(body) >= 26:       # WAVE_FORMAT_EXTENSIBLE: sub-format tag
                fmt = (struct.unpack("<H", body[24:26])[0],) + fmt[1:]
        elif cid == b"data":
            raw = body
        pos += 8 + size + (size & 1)
    if fmt is None or raw is None:
        raise ValueError(f"{file_path}: missing fmt/data chunk")
    return fmt, raw


def _read_wav_pcm16(file_path):
    """Raw frames of a 16-bit PCM mono/stereo WAV file: (int16 [frames * channels], channels, rate), else None."""
    fmt, raw = _wav_chunks(file_path)
    tag, channels, rate, _, _, bits = fmt
    if tag != 1 or bits != 16 or channels not in (1, 2):
        return None
    x = np.frombuffer(raw, dtype="<i2")
    return x[:len(x) // channels * channels], int(channels), int(rate)


def _read_wav(file_path):
    """Minimal RIFF/WAVE reader (PCM 8/16/24/32-bit, IEEE float 32/64) -> (float32 [frames, channels], rate)."""
    fmt, raw = _wav_chunks(file_path)
    tag, channels, rate, _, _, bits = fmt
    if tag == 1:
        if bits == 8:
            x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
        elif bits == 16:
            x = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
        elif bits == 24:
            b = np.frombuffer(raw[:len(raw) // 3 * 3], dtype=np.uint8).reshape(-1, 3).astype(np.int32)
            v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
            v = np.where(v & 0x800000, v - 0x1000000, v)
            x = v.astype(np.float32) / 8388608.0
        elif bits == 32:
            x = (np.frombuffer(raw, dtype="<i4").astype(np.float64) / 2147483648.0).astype(np.float32)
        else:
            raise UnsupportedAudioFormat(f"{file_path}: unsupported PCM width {bits}")
    elif tag == 3:
        x = np.frombuffer(raw, dtype="<f4" if bits == 32 else "<f8").astype(np.float32)
    else:
        raise UnsupportedAudioFormat(f"{file_path}: unsupported WAVE format tag {tag}")
    x = x[:len(x) // channels * channels].reshape(-1, channels)
    return x, int(rate)


# --------------------------------------------------------------------------- the reference's five functions
def load_audio(file_path, sr=Config.SAMPLE_RATE, duration=Config.AUDIO_DURATION):
    """reference :12-19 -- (float32[sr*duration], sr): first `duration` seconds, mono, zero-padded / trimmed.

    Files already at `sr` Hz take the exact path of the reference (decode -> mono mean -> pad/trim).  Other
    rates are resampled with a polyphase Kaiser filter (scipy.signal.resample_poly), which is NOT librosa's
    soxr_hq resampler: that front-end is row f3 ("next") of the scope table and carries no parity claim.
    """
    try:
        x, native = _read_wav(file_path)
    except UnsupportedAudioFormat as e:                    # mp3 / ogg / compressed WAVE: the host decoders of librosa.load
        x, native = _host_decode(file_path, str(e).split(": ", 1)[-1])
    if duration is not None:
        x = x[:int(round(native * duration))]
    audio = x.mean(axis=1, dtype=np.float32) if x.shape[1] > 1 else x[:, 0]
    if native != sr:
        from math import gcd
        from scipy.signal import resample_poly
        g = gcd(int(sr), int(native))
        audio = resample_poly(audio.astype(np.float64), int(sr) // g, int(native) // g).astype(np.float32)
    fixed = np.zeros(sr * duration, dtype=np.float32)          # zero tail for short files, cut for long ones
    keep = min(len(audio), len(fixed))
    fixed[:keep] = audio[:keep]
    return fixed, sr


def extract_mfcc(audio, sr, n_mfcc=Config.N_MFCC):
    """reference :22-24 -- frame-mean of librosa.feature.mfcc(y, sr, n_mfcc): shape (n_mfcc,)."""
    row, dt = _features_1clip(audio, sr, n_mfcc)
    return row[:n_mfcc].astype(dt)


def extract_chroma(audio, sr):
    """reference :27-29 -- frame-mean of librosa.feature.chroma_stft(y, sr): shape (12,)."""
    row, dt = _features_1clip(audio, sr, Config.N_MFCC)
    return row[Config.N_MFCC:Config.N_MFCC + 12].astype(dt)


def extract_spectral_features(audio, sr):
    """reference :32-37 -- np.float32[4] = [zcr, spectral_centroid, spectral_rolloff, rms] frame means."""
    row, _ = _features_1clip(audio, sr, Config.N_MFCC)
    return np.array(row[Config.N_MFCC + 12:Config.N_MFCC + 16], dtype=np.float32)     # zcr, centroid, rolloff, rms


def preprocess_audio(file_path):
    """reference :40-46 -- file -> np.float32[56] = [40 mfcc | 12 chroma | zcr, centroid, rolloff, rms].

    A 16-bit PCM WAV file is handed to the device as raw frames (what load_audio does happens there, bit-identically,
    see preprocess_audio_batch), so a 48 kHz file costs one device pass instead of a host resample_poly."""
    try:
        raw = _read_wav_pcm16(file_path)
    except ValueError:
        raw = None                      # not a RIFF file: let load_audio raise its own error below
    if raw is not None:
        return preprocess_audio_batch([file_path])[0]
    wave_, rate = load_audio(file_path)
    parts = (extract_mfcc(wave_, rate), extract_chroma(wave_, rate), extract_spectral_features(wave_, rate))
    return np.concatenate(parts).astype(np.float32)


# --------------------------------------------------------------------------- additive batched entry points
def extract_features_batch(waveforms, lengths=None, sr=Config.SAMPLE_RATE, n_mfcc=Config.N_MFCC, n_samples=None):
    """[B, L] float32 (cuda tensor -> cuda tensor, host array -> host array) -> [B, n_mfcc + 16].
    `lengths` (int32 [B]) gives per-clip sample counts for padded batches; rows of clips that produced
    non-finite features (non-finite samples) raise ParameterError on the host path."""
    from sfx_b200 import extract_features_batch as _batch
    out = _batch(waveforms, lengths, sr=int(sr), n_mfcc=int(n_mfcc), n_samples=n_samples)
    if isinstance(out, np.ndarray) and not np.isfinite(out).all():
        bad = np.nonzero(~np.isfinite(out).all(axis=1))[0]
        raise ParameterError(f"Audio buffer is not finite everywhere (clips {bad[:8].tolist()}...)")
    return out


def preprocess_audio_batch(file_paths, on_error="raise"):
    """Batched preprocess_audio: float32 [N, 56] for N files, one device pass per (sample rate, channel count) group.

    16-bit PCM WAV files (mono / stereo, any rate) never touch the host's float path: their raw frames go to the device,
    which does what load_audio does -- x/32768, channel mean, polyphase resampling to Config.SAMPLE_RATE, pad / trim --
    and then extracts the features (sfx_preprocess_host_pcm16; bit-identical to load_audio + extract on the host path).
    Other encodings are decoded by load_audio on the host and extracted in one batch.
    on_error='skip' mirrors the per-file try/except of train_speech_model.py:124,142-143 and returns
    (features, kept_indices); on_error='collect' returns (features [N, 56] with NaN rows for failed files, {index: error})."""
    file_paths = list(file_paths)
    errors = {}
    sr, duration = Config.SAMPLE_RATE, Config.AUDIO_DURATION
    rows = {}                                    # index -> float32[56]
    groups, host_clips, host_idx = {}, [], []
    for i, fp in enumerate(file_paths):
        try:
            raw = _read_wav_pcm16(fp)
            if raw is not None:
                x, channels, rate = raw
                if len(x) == 0:
                    raise ParameterError("Audio data must be a non-empty 1-D array (mono)")
                groups.setdefault((rate, channels), []).append((i, x))
            else:
                audio, _ = load_audio(fp)
                _valid_audio(audio)
                host_clips.append(audio)
                host_idx.append(i)
        except Exception as e:  # noqa: BLE001  (reference :142 catches everything)
            if on_error not in ("skip", "collect"):
                raise
            errors[i] = e
    if groups:
        from sfx_b200 import get_extractor
        ex = get_extractor(None, int(sr))
        for (rate, channels), items in groups.items():
            limit = int(round(rate * duration)) * channels
            width = min(limit, max(len(x) for _, x in items))
            pcm = np.zeros((len(items), width + (width & 1)), dtype=np.int16)
            frames = np.zeros(len(items), dtype=np.int32)
            for r, (_, x) in enumerate(items):
                n = min(len(x), width)
                pcm[r, :n] = x[:n]
                frames[r] = n // channels
            feats = ex.preprocess_pcm16(pcm, frames, rate, channels=channels, duration=duration)
            for r, (i, _) in enumerate(items):
                rows[i] = feats[r]
    if host_clips:
        feats = extract_features_batch(np.stack(host_clips))
        for r, i in enumerate(host_idx):
            rows[i] = feats[r]
    kept = sorted(rows)
    if on_error == "collect":
        full = np.full((len(file_paths), 56), np.nan, dtype=np.float32)
        for i in kept:
            full[i] = rows[i]
        return full, errors
    out = np.stack([rows[i] for i in kept]).astype(np.float32) if kept else np.zeros((0, 56), dtype=np.float32)
    return (out, kept) if on_error == "skip" else out
