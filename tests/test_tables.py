"""Host-side tables (sfx_b200/tables.py) against the oracle's filterbanks, and the numpy model of the kernel's
per-warp FFT against numpy's rfft."""
import numpy as np
import scipy.fftpack

import fft_model
from oracle import librosa_port as lp
from sfx_b200 import tables


def test_banks_bit_identical_to_oracle():
    tb = tables.build_tables(22050)
    assert np.array_equal(tb["mel_dense"], lp.mel_filterbank())
    e = tb["edges"]
    assert np.array_equal(e, np.linspace(-0.5, 0.5, 101)) and e[50] == 0.0
    for i in (0, 17, 50, 83, 99):
        ref = lp.chroma_filterbank(tuning=float(e[i]))
        assert np.array_equal(tb["chroma_f32"][i, :, :1025], ref)
        # device operand: hi + 2^-11 * lo in float16 reproduces the bank to ~2^-21 relative (+ fp16 subnormal floor)
        hi, lo = tb["chroma16"][i, 0].astype(np.float64), tb["chroma16"][i, 1].astype(np.float64)
        rec = hi + lo / 2048.0
        assert np.all(np.abs(rec[:, :1025] - ref) <= np.abs(ref) * 2.0 ** -20 + 2.0 ** -34)
        assert not tb["chroma16"][i, :, :, 1025:].any()
        assert np.array_equal(tb["chroma_ny"][i], ref[:, 1024])
    assert np.array_equal(tb["hann"], lp.hann_window().astype(np.float32))
    assert (tb["pip_kmin"], tb["pip_kmax"]) == (14, 371)


def test_dct_matrix():
    tb = tables.build_tables(22050)
    x = np.random.default_rng(0).standard_normal((128, 3))
    assert np.abs(tb["dct"] @ x - scipy.fftpack.dct(x, axis=0, type=2, norm="ortho")).max() < 1e-12


def test_mel_sparse_layout_roundtrip():
    tb = tables.build_tables(22050)
    dense = np.zeros((128, 1025 + 64), dtype=np.float32)
    for m in range(128):
        s, lane = divmod(m, 32)
        for i in range(tb["mel_len"][s]):
            dense[m, tb["mel_lo"][m] + i] += tb["melw"][tb["mel_off"][s] + i, lane]
    assert np.array_equal(dense[:, :1025], tb["mel_dense"]) and not dense[:, 1025:].any()
    assert int((tb["mel_dense"] > 0).sum()) == 2018


def test_other_sample_rate_tables():
    tb = tables.build_tables(16000)
    assert np.array_equal(tb["mel_dense"], lp.mel_filterbank(sr=16000))
    assert np.array_equal(tb["chroma_f32"][50, :, :1025], lp.chroma_filterbank(sr=16000, tuning=0.0))


def test_warp_fft_model_matches_rfft():
    tb = tables.build_tables(22050)
    rng = np.random.default_rng(1)
    for _ in range(3):
        x = rng.standard_normal(2048)
        X = fft_model.warp_rfft2048(x, tb["tw1"].astype(np.float64), tb["tw2"].astype(np.float64))
        ref = np.fft.rfft(x)
        assert np.abs(X - ref).max() / np.abs(ref).max() < 2e-7
