"""Scope row f3: the device front-end (16-bit PCM -> x/32768 -> mono -> polyphase resampling -> pad/trim -> features).

CPU part: the restated resample_poly filter / summation (sfx_b200/resample.py) is bit-identical to scipy.signal.resample_poly.
GPU part: sfx_preprocess_host_pcm16 equals, bit for bit, the host path of load_audio followed by the extractor."""
import os
import sys
import wave

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "multimodal-emotion-classification_b200"), ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import synth
from sfx_b200 import resample

SR, DUR = 22050, 3
RATES = [(48000, 1), (24414, 1), (44100, 2), (16000, 2), (22050, 2), (8000, 1)]


@pytest.mark.parametrize("native", [48000, 24414, 44100, 16000, 8000, 32000])
def test_restated_resample_poly_is_bit_identical_to_scipy(native):
    from scipy.signal import resample_poly
    rng = np.random.default_rng(native)
    flt = resample.resample_filter(native, SR)
    for n in (1, 7, 400, 1501):
        x = (rng.integers(-32768, 32768, n).astype(np.float32) / np.float32(32768)).astype(np.float64)
        ref = resample_poly(x, flt["up"], flt["down"])
        got = resample.resample_poly_direct(x, flt)
        assert got.shape == ref.shape and np.array_equal(got, ref)


def test_native_rate_filter_is_identity():
    flt = resample.resample_filter(SR, SR)
    assert flt["up"] == flt["down"] == 1
    x = np.arange(5, dtype=np.float64)
    assert np.array_equal(resample.resample_poly_direct(x, flt), x)


def host_load_audio(pcm_row, frames, native, channels, DUR=DUR):
    """What preprocessing.audio_preprocessing.load_audio does to a decoded PCM16 file (reference :12-19 + resample_poly)."""
    from scipy.signal import resample_poly
    x = pcm_row[:frames * channels].astype(np.float32) / np.float32(32768.0)
    x = x.reshape(-1, channels)[:int(round(native * DUR))]
    audio = x.mean(axis=1, dtype=np.float32) if channels > 1 else x[:, 0]
    if native != SR:
        flt = resample.resample_filter(native, SR)
        audio = resample_poly(audio.astype(np.float64), flt["up"], flt["down"]).astype(np.float32)
    out = np.zeros(SR * DUR, dtype=np.float32)
    n = min(len(audio), SR * DUR)
    out[:n] = audio[:n]
    return out


@pytest.fixture(scope="module")
def ex():
    from sfx_b200 import get_extractor
    return get_extractor(torch.device("cuda", 0))


@pytest.mark.gpu
@pytest.mark.parametrize("native,channels", RATES)
def test_device_front_end_equals_host_load_audio(ex, native, channels):
    rng = np.random.default_rng(native + channels)
    B = 6
    secs = [3.4, 3.0, 2.2, 0.7, 0.05, 1.9]                       # longer than, equal to and shorter than the 3 s window
    frames = np.array([int(native * s) for s in secs], dtype=np.int32)
    L = int(frames.max()) * channels
    pcm = np.zeros((B, L + (L & 1)), dtype=np.int16)
    for i in range(B):
        n = int(frames[i])
        kinds = synth.KINDS
        y = np.stack([synth.make_clip(kinds[(i + c) % len(kinds)], n, rng) for c in range(channels)], axis=1)
        pcm[i, :n * channels] = np.clip(np.round(y * 32767.0), -32768, 32767).astype(np.int16).reshape(-1)
    got = ex.preprocess_pcm16(pcm, frames, native, channels=channels, duration=DUR)
    waves = np.stack([host_load_audio(pcm[i], int(frames[i]), native, channels) for i in range(B)])
    ref = ex.extract_host(waves)
    assert np.isfinite(got).all()
    assert np.array_equal(got, ref), np.abs(got - ref).max(axis=0)
    # small chunks / pageable vs pinned input give the same rows
    assert np.array_equal(ex.preprocess_pcm16(pcm, frames, native, channels=channels, duration=DUR, chunk_clips=4), ref)
    pinned = torch.from_numpy(pcm).pin_memory()
    assert np.array_equal(ex.preprocess_pcm16(pinned.numpy(), frames, native, channels=channels, duration=DUR), ref)


@pytest.mark.gpu
def test_device_front_end_long_clip_uses_64_bit_indices(ex):
    """30 s at 24 414 Hz: frames * up exceeds 2^31, so the kernel switches to 64-bit index arithmetic; same bits as the host."""
    native, dur = 24414, 30
    rng = np.random.default_rng(9)
    frames = np.array([native * dur + 500, native * 11], dtype=np.int32)
    pcm = np.zeros((2, int(frames.max())), dtype=np.int16)
    for i in range(2):
        y = synth.make_clip(synth.KINDS[i], int(frames[i]), rng)
        pcm[i, :frames[i]] = np.clip(np.round(y * 32767.0), -32768, 32767).astype(np.int16)
    got = ex.preprocess_pcm16(pcm, frames, native, duration=dur)
    waves = np.stack([host_load_audio(pcm[i], int(frames[i]), native, 1, DUR=dur) for i in range(2)])
    assert np.array_equal(got, ex.extract_host(waves))


@pytest.mark.gpu
def test_preprocess_audio_batch_uses_the_device_front_end(ex, tmp_path):
    """RAVDESS-like (48 kHz mono) and TESS-like (24 414 Hz) 16-bit files plus a stereo 44.1 kHz one: the batched call must
    give exactly what preprocess_audio gives file by file (host decode + host resample_poly + extractor)."""
    from preprocessing.audio_preprocessing import preprocess_audio, preprocess_audio_batch
    rng = np.random.default_rng(11)
    paths = []
    for k, (rate, ch, secs) in enumerate([(48000, 1, 3.6), (24414, 1, 2.1), (44100, 2, 1.3), (22050, 1, 3.0), (48000, 1, 0.9)]):
        n = int(rate * secs)
        y = np.stack([synth.make_clip(synth.KINDS[(k + c) % 4], n, rng) for c in range(ch)], axis=1)
        p = os.path.join(tmp_path, f"f{k}.wav")
        with wave.open(p, "wb") as wf:
            wf.setnchannels(ch); wf.setsampwidth(2); wf.setframerate(rate)
            wf.writeframes(np.clip(np.round(y * 32767.0), -32768, 32767).astype("<i2").tobytes())
        paths.append(p)
    from preprocessing.audio_preprocessing import extract_chroma, extract_mfcc, extract_spectral_features, load_audio
    batch = preprocess_audio_batch(paths)

    def host_path(p):                 # the reference's own composition (:40-46) on the host-decoded, host-resampled audio
        audio, sr = load_audio(p)
        return np.concatenate([extract_mfcc(audio, sr), extract_chroma(audio, sr), extract_spectral_features(audio, sr)])

    host = np.stack([host_path(p) for p in paths]).astype(np.float32)
    assert batch.shape == (5, 56) and np.array_equal(batch, host)
    assert np.array_equal(np.stack([preprocess_audio(p) for p in paths]), host)
