"""CPU tests of the oracle (oracle/librosa_port.py): committed golden vectors, hand-derived known answers and
independent cross-checks (torchaudio's librosa-compatible MFCC front-end, transformers' librosa-derived chroma
bank).  The reference's own tests pin only shapes/finiteness (tests/test_preprocessing.py:30-67); those are
mirrored at the bottom."""
import os
import zlib

import numpy as np
import pytest

import synth
from oracle import librosa_port as lp

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SR = 22050


def test_golden_config1_subset():
    g = np.load(os.path.join(GOLD, "config1_seed0.npz"))
    w = synth.make_batch(64, 66150, seed=0)
    assert np.uint32(zlib.crc32(w.tobytes())) == g["crc"], "synthetic generator drifted from the fixtures"
    idx = [0, 1, 2, 3, 17, 42]                      # a few clips keep the CPU suite fast
    got = lp.features_batch(w[idx])
    np.testing.assert_allclose(got, g["features"][idx], rtol=1e-6, atol=1e-6)
    for i in idx[:4]:
        assert lp.debug_intermediates(w[i])["tuning"] == pytest.approx(g["tuning"][i], abs=1e-12)


def test_golden_ragged_and_edges():
    g = np.load(os.path.join(GOLD, "ragged_seed3.npz"))
    wr, lens = synth.make_ragged(10, 11025, 200000, seed=3)
    assert np.array_equal(lens, g["lengths"])
    got = lp.features_batch(wr[:4], lens[:4])
    np.testing.assert_allclose(got, g["features"][:4], rtol=1e-6, atol=1e-6)
    e = np.load(os.path.join(GOLD, "edge_cases.npz"))
    rng = np.random.default_rng(5)
    clips = np.stack([synth.make_clip(k, 66150, rng) for k in ("zero", "dc", "square")])
    np.testing.assert_allclose(lp.features_batch(clips), e["features"], rtol=1e-6, atol=1e-6)


def test_known_answer_silence():
    f = lp.features_from_audio(np.zeros(66150, dtype=np.float32))
    assert f[0] == pytest.approx(-100.0 * np.sqrt(128.0), rel=1e-6)       # every log-mel bin = -100 dB
    assert np.abs(f[1:40]).max() < 1e-3
    assert np.all(f[40:] == 0.0)                                          # chroma, zcr, centroid, rolloff, rms


def test_known_answer_bin_centred_sinusoid():
    k = 93                                                                # ~1001 Hz, exactly on an rFFT bin
    f0 = k * SR / 2048.0
    t = np.arange(66150) / SR
    y = (0.5 * np.sin(2 * np.pi * f0 * t)).astype(np.float32)
    roll = lp.spectral_rolloff(y, SR)[0, 4:-4]                            # interior frames: pure tone
    assert np.all(np.abs(roll - f0) <= SR / 2048.0 * 1.01)                # roll-off within a bin of the tone
    cent = lp.spectral_centroid(y, SR)[0, 4:-4]
    assert np.all(np.abs(cent - f0) < 1.0)                                # Hann main lobe is symmetric
    assert np.allclose(lp.rms(y)[0, 4:-4], 0.5 / np.sqrt(2), rtol=2e-3)
    assert np.allclose(lp.zero_crossing_rate(y)[0, 4:-4], 2 * f0 / SR, rtol=0.02)   # two crossings per period
    tuning = lp.debug_intermediates(y)["tuning"]
    res = np.mod(12 * np.log2(f0 / 27.5), 1.0)
    res = res - 1 if res >= 0.5 else res
    assert abs(tuning - res) <= 0.011
    chroma = lp.extract_chroma(y, SR)
    assert chroma.argmax() == int(np.round(12 * np.log2(f0 / 27.5) - tuning + 9)) % 12   # bank row 0 = C


def test_known_answer_dc():
    y = np.full(66150, 0.25, dtype=np.float32)
    f = lp.features_from_audio(y)
    assert f[52] == 0.0                                                   # no sign changes
    assert f[55] == pytest.approx(0.25, rel=0.02)
    assert f[53] < 50.0 and f[54] < 200.0                                 # energy in the lowest bins (edge frames leak)


def test_mfcc_against_torchaudio():
    torch = pytest.importorskip("torch")
    torchaudio = pytest.importorskip("torchaudio")
    y = synth.make_batch(2, 66150, seed=7)
    tr = torchaudio.transforms.MFCC(sample_rate=SR, n_mfcc=40, dct_type=2, norm="ortho", log_mels=False,
                                    melkwargs=dict(n_fft=2048, hop_length=512, n_mels=128, center=True,
                                                   pad_mode="constant", power=2.0, norm="slaney", mel_scale="slaney",
                                                   f_min=0.0, f_max=SR / 2))
    for clip in y:
        ta = tr(torch.from_numpy(clip)).numpy()
        mine = lp.mfcc(clip)
        assert np.abs(ta - mine).max() < 2e-3 * max(1.0, np.abs(mine).max() / 100)


def test_filterbanks_against_independent_ports():
    torchaudio = pytest.importorskip("torchaudio")
    au = pytest.importorskip("transformers.audio_utils")
    fb = torchaudio.functional.melscale_fbanks(1025, 0.0, SR / 2, 128, SR, norm="slaney", mel_scale="slaney").numpy().T
    assert np.abs(fb - lp.mel_filterbank()).max() < 1e-6
    for tun in (0.0, -0.33, 0.49):
        ref = au.chroma_filter_bank(num_frequency_bins=2048, num_chroma=12, sampling_rate=SR, tuning=tun, power=2,
                                    weighting_parameters=(5.0, 2.0), start_at_c_chroma=True).astype(np.float32)
        assert np.array_equal(ref, lp.chroma_filterbank(tuning=tun))
    x = np.random.default_rng(0).random((128, 5)).astype(np.float32) * 3
    assert np.allclose(lp.power_to_db(x), au.power_to_db(x, reference=1.0, min_value=1e-10, db_range=80.0), atol=1e-5)


def test_stft_and_centroid_against_torch():
    """Independent implementations of the framing / FFT and of the centroid formula: torch.stft with the same centring and
    zero padding, and torchaudio's spectral_centroid (reflect-padded, so only interior frames are comparable)."""
    torch = pytest.importorskip("torch")
    torchaudio = pytest.importorskip("torchaudio")
    y = synth.make_batch(2, 40000, seed=9)
    win = torch.hann_window(2048, periodic=True, dtype=torch.float64)
    for clip in y:
        X = lp.stft(clip)
        Xt = torch.stft(torch.from_numpy(clip).double(), n_fft=2048, hop_length=512, win_length=2048, window=win, center=True,
                        pad_mode="constant", return_complex=True).numpy()
        assert X.shape == Xt.shape
        assert np.abs(X - Xt).max() < 2e-6 * np.abs(Xt).max()
        c = lp.spectral_centroid(clip)[0]
        ct = torchaudio.functional.spectral_centroid(torch.from_numpy(clip), SR, pad=0, window=win.float(), n_fft=2048,
                                                     hop_length=512, win_length=2048).numpy()
        assert c.shape == ct.shape
        inner = slice(3, len(c) - 3)
        assert np.abs(c[inner] - ct[inner]).max() < 1e-3 * np.abs(c[inner]).max()


def test_frame_count_and_dtype_trail():
    for n in (11025, 66150, 81585):
        X = lp.stft(np.zeros(n, dtype=np.float32))
        assert X.shape == (1025, 1 + n // 512) and X.dtype == np.complex64
    assert lp.stft(np.zeros(4096, dtype=np.float64)).dtype == np.complex128


def test_invalid_audio_raises():
    with pytest.raises(lp.ParameterError):
        lp.extract_mfcc(np.array([0.0, np.nan, 0.0], dtype=np.float32))
    with pytest.raises(lp.ParameterError):
        lp.extract_mfcc(np.zeros(10, dtype=np.int16))


class TestAudioPreprocessingShapes:
    """The reference's tests/test_preprocessing.py:30-67, run against the oracle (float64 randn input)."""
    audio = np.random.default_rng(0).standard_normal(SR * 3)

    def test_mfcc_extraction(self):
        m = lp.extract_mfcc(self.audio, SR)
        assert m.shape == (40,) and np.all(np.isfinite(m))

    def test_chroma_extraction(self):
        c = lp.extract_chroma(self.audio, SR)
        assert c.shape == (12,) and np.all(np.isfinite(c))

    def test_spectral_features(self):
        s = lp.extract_spectral_features(self.audio, SR)
        assert s.shape == (4,) and np.all(np.isfinite(s))
