"""GPU parity tests (pytest -m gpu): the sm_100a path, called through the C ABI (ctypes -> libsfx_b200.so),
against the CPU oracle on the same seeded inputs, the committed golden fixtures, and size-independent
properties at full batch sizes.  Tolerance: |err| <= 1e-3*|ref| + atol(group), see tests/synth.py."""
import os
import zlib

import numpy as np
import pytest
import torch

import synth
from oracle import librosa_port as lp

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
N3S = 66150


@pytest.fixture(scope="module")
def ex():
    from sfx_b200 import get_extractor
    return get_extractor(torch.device("cuda", 0))


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def assert_parity(got, ref, n_mfcc=40):
    ok, report = synth.compare(got, ref, n_mfcc=n_mfcc)
    assert ok, "\n" + report


def test_native_library_is_loaded(ex):
    maps = open("/proc/self/maps").read()
    assert "libsfx_b200.so" in maps
    assert ex.lib.sfx_device_count() >= 1 and ex.lib.sfx_launches_per_extract() >= 1


def test_fp32_peak_helper_measures_a_plausible_rate(ex):
    """bench.py's compute denominator: a B200 has 148 SMs x 128 FP32 lanes x 2 flop at <= 2.1 GHz = 79.6 TFLOP/s."""
    import ctypes
    tf = ctypes.c_double(0.0)
    from sfx_b200 import _lib
    assert _lib.load_bench().sfx_measure_fp32_peak(ex.index, ctypes.byref(tf)) == 0
    assert 40.0 < tf.value < 80.0


def test_config1_golden_64_clips(ex):
    """BASELINE configs[0]: 64 synthetic 3 s clips, seed 0, against the committed oracle fixture."""
    g = np.load(os.path.join(GOLD, "config1_seed0.npz"))
    w = synth.make_batch(64, N3S, seed=0)
    assert np.uint32(zlib.crc32(w.tobytes())) == g["crc"]
    dbg = {}
    got = ex.extract(dev(w), debug=dbg).cpu().numpy()
    assert_parity(got, g["features"])
    tuning = dbg["clip_info"].cpu().numpy()[:, 0]
    flips = int((np.abs(tuning - g["tuning"]) > 1e-6).sum())
    assert flips == 0, f"{flips} tuning arg-max flips vs the oracle"
    got2 = ex.extract(dev(w)).cpu().numpy()                   # non-debug kernel instantiation, bitwise equal
    assert np.array_equal(got, got2)


def test_edge_cases_golden(ex):
    e = np.load(os.path.join(GOLD, "edge_cases.npz"))
    rng = np.random.default_rng(5)
    clips = np.stack([synth.make_clip(k, N3S, rng) for k in ("zero", "dc", "square")])
    got = ex.extract(dev(clips)).cpu().numpy()
    assert_parity(got, e["features"])
    assert got[0, 0] == pytest.approx(-100.0 * np.sqrt(128.0), rel=1e-6)       # silence: every band -100 dB
    assert np.abs(got[0, 1:40]).max() < 1e-3 and not got[0, 40:].any()


def test_ragged_golden_and_poisoned_padding(ex):
    g = np.load(os.path.join(GOLD, "ragged_seed3.npz"))
    wr, lens = synth.make_ragged(10, 11025, 200000, seed=3)
    got = ex.extract(dev(wr), dev(lens)).cpu().numpy()
    assert_parity(got, g["features"])


def test_config2_ravdess_shaped(ex):
    """BASELINE configs[1] shape: ~3.7 s clips; preprocess_audio semantics (truncate to 3 s, T=130) and raw (T=160)."""
    rng = np.random.default_rng(11)
    n_raw = 81586
    w = synth.make_batch(1440, n_raw, seed=11)
    d = dev(w)
    trunc = ex.extract(d, n_samples=N3S).cpu().numpy()
    raw = ex.extract(d).cpu().numpy()
    assert np.isfinite(trunc).all() and np.isfinite(raw).all()
    idx = rng.choice(1440, size=24, replace=False)
    assert_parity(trunc[idx], lp.features_batch(w[idx, :N3S]))
    assert_parity(raw[idx[:8]], lp.features_batch(w[idx[:8]]))
    # truncation == extracting the first 66150 samples laid out contiguously, bit for bit
    again = ex.extract(dev(w[:64, :N3S])).cpu().numpy()
    assert np.array_equal(again, trunc[:64])


def test_config3_tess_shaped_zero_tail(ex):
    """BASELINE configs[2] shape: ~2 s of signal zero-padded to 3 s by load_audio -> top_db clamp is active."""
    rng = np.random.default_rng(12)
    w = np.zeros((256, N3S), dtype=np.float32)
    for i in range(256):
        n = int(44100 * rng.uniform(0.85, 1.15))
        w[i, :n] = synth.make_clip(synth.KINDS[i % 2], n, rng)
    got = ex.extract(dev(w)).cpu().numpy()
    idx = rng.choice(256, size=16, replace=False)
    assert_parity(got[idx], lp.features_batch(w[idx]))


def test_config2_full_batch_oracle_parity(ex):
    """BASELINE configs[1] at its named size: all 1 440 ~3.7 s clips against the oracle (16-core process pool), both as
    preprocess_audio sees them (truncated to 3 s, T = 130) and raw (T = 160)."""
    import oracle_pool
    n_raw = 81586
    w, (ref_trunc, ref_raw) = oracle_pool.indexed_batch(1440, n_raw, seed=211, n_used=[N3S, n_raw])
    d = dev(w)
    assert_parity(ex.extract(d, n_samples=N3S).cpu().numpy(), ref_trunc)
    assert_parity(ex.extract(d).cpu().numpy(), ref_raw)


def test_config3_full_batch_oracle_parity(ex):
    """BASELINE configs[2] at its named size: all 2 800 ~2 s clips zero-padded to 3 s (top_db clamp active on every clip)."""
    import oracle_pool
    B = 2800
    rng = np.random.default_rng(212)
    w, _ = oracle_pool.indexed_batch(B, N3S, seed=212, kinds=synth.KINDS[:2], n_used=[])        # the clips only
    cut = (44100 * rng.uniform(0.85, 1.15, B)).astype(np.int64)
    w[np.arange(N3S)[None, :] >= cut[:, None]] = 0.0
    got = ex.extract(dev(w)).cpu().numpy()
    assert_parity(got, oracle_pool.oracle_rows(w))


def test_config5_at_size_4096_clips_half_to_sixty_seconds(ex):
    """BASELINE configs[4] at its named size: 4 096 clips, lengths log-uniform in [0.5 s, 60 s] (T = 22 .. 2 584), padded rows
    + lengths, poisoned padding.  The batch (21.7 GB) is synthesised on the device; oracle parity on 72 sampled clips that
    include the longest and the shortest; the host entry point on a 384-clip slice must give the device path's rows."""
    import oracle_pool
    B, n_min, n_max = 4096, 11025, 1323000
    rng = np.random.default_rng(213)
    lens = np.exp(rng.uniform(np.log(n_min), np.log(n_max), size=B)).astype(np.int64)
    lens[0], lens[-1], lens[B // 2] = n_min, n_max, n_max - 511
    lens = lens.astype(np.int32)
    g = torch.Generator(device="cuda").manual_seed(213)
    w = torch.empty((B, n_max), dtype=torch.float32, device="cuda")
    t = torch.arange(n_max, device="cuda", dtype=torch.float32) / 22050.0
    for c0 in range(0, B, 256):                                       # noise rows; every 2nd row a harmonic stack on top
        blk = w[c0:c0 + 256]
        blk.normal_(0.0, 0.1, generator=g)
        f0 = 90.0 + 210.0 * torch.rand((128, 1), device="cuda", generator=g)
        y = torch.zeros((128, n_max), device="cuda")
        for h in range(1, 9):
            y += torch.sin(6.2831853 * h * f0 * t) / h
        blk[1::2] = 0.3 * y / y.abs().amax(dim=1, keepdim=True) + 0.002 * blk[1::2]
        del y
    ld = torch.from_numpy(lens).cuda()
    w.masked_fill_(torch.arange(n_max, device="cuda")[None, :] >= ld[:, None], 7.0)            # poison: must never be read
    got = ex.extract(w, ld)
    assert ex.lib.sfx_launches_per_extract() == 2                     # longest-first order kernel + extractor
    assert bool(torch.isfinite(got).all())
    idx = np.unique(np.concatenate([[0, B - 1, B // 2, int(np.argmax(lens)), int(np.argmin(lens))],
                                    rng.choice(B, size=67, replace=False)]))
    wl = w[torch.from_numpy(idx).cuda()].cpu().numpy()
    ref = oracle_pool.oracle_rows(wl, lens[idx])
    assert_parity(got[torch.from_numpy(idx).cuda()].cpu().numpy(), ref)
    sl = slice(B // 2 - 192, B // 2 + 192)                            # includes a 60 s clip
    host = ex.extract_host(w[sl].cpu().numpy(), lens[sl])
    assert np.array_equal(host, got[sl].cpu().numpy())


def test_two_threads_two_sample_rates(ex):
    """The C ABI from two host threads at once: one extracts at 22 050 Hz in a loop while the other uploads the table sets
    of new sample rates (which grows the per-device table list) and extracts with them; rows must equal the serial ones."""
    import threading
    from sfx_b200.extractor import SpeechFeatureExtractor
    w22 = dev(synth.make_batch(64, N3S, seed=71))
    ref22 = ex.extract(w22).clone()
    rates = (11025, 32000, 12000)
    waves = {sr: dev(synth.make_batch(8, 2 * sr, seed=72)) for sr in rates}
    errors, results = [], {}

    def worker_a():
        try:
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                for _ in range(40):
                    if not torch.equal(ex.extract(w22), ref22):
                        errors.append("22 050 Hz rows changed while another thread initialised tables")
            s.synchronize()
        except Exception as e:      # noqa: BLE001
            errors.append(repr(e))

    def worker_b():
        try:
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                for sr in rates:
                    e2 = SpeechFeatureExtractor(torch.device("cuda", 0), sr=sr)
                    results[sr] = e2.extract(waves[sr]).clone()
            s.synchronize()
        except Exception as e:      # noqa: BLE001
            errors.append(repr(e))

    ta, tb_ = threading.Thread(target=worker_a), threading.Thread(target=worker_b)
    ta.start(); tb_.start(); ta.join(); tb_.join()
    assert not errors, errors
    from sfx_b200 import get_extractor
    for sr in rates:
        serial = get_extractor(torch.device("cuda", 0), sr=sr).extract(waves[sr])
        assert torch.equal(results[sr], serial)
        assert_parity(serial[:2].cpu().numpy(), np.stack([lp.features_from_audio(x, sr=sr) for x in waves[sr][:2].cpu().numpy()]))


def test_host_path_reports_non_finite_clips(ex):
    """include/sfx.h: a clip with a NaN / Inf sample gets a NaN row and the host entry point returns SFX_ERR_BAD_CLIP after
    delivering every row; a bad length is rejected before any work."""
    w = synth.make_batch(9, N3S, seed=73)
    w[4, 30000] = np.nan
    w[7, 100] = np.inf
    out = np.zeros((9, 56), dtype=np.float32)
    rc = ex.lib.sfx_extract_host(ex.index, 22050, w.ctypes.data, N3S, None, N3S, 9, 40, out.ctypes.data, 56, 0)
    assert rc == -5 and b"clip 4" in ex.lib.sfx_last_error()
    good = [0, 1, 2, 3, 5, 6, 8]
    assert np.isnan(out[[4, 7]]).any(axis=1).all() and np.isfinite(out[good]).all()
    assert np.array_equal(out[good], ex.extract(dev(w[good])).cpu().numpy())
    rows = ex.extract_host(w)                                             # the Python wrapper returns the rows
    assert np.array_equal(np.isnan(rows).any(axis=1), np.isnan(out).any(axis=1))
    lens = np.array([N3S] * 8 + [0], dtype=np.int32)
    out2 = np.zeros_like(out)
    rc = ex.lib.sfx_extract_host(ex.index, 22050, w.ctypes.data, N3S, lens.ctypes.data, N3S, 9, 40, out2.ctypes.data, 56, 0)
    assert rc == -5 and not out2.any()


@pytest.mark.parametrize("mode", ["fused", "fused_umma", "stream", "split"])
def test_every_pipeline_against_the_oracle(ex, mode):
    """All three schedules of the one arithmetic, forced, on a uniform and a ragged batch (incl. invalid lengths)."""
    w = synth.make_batch(40, N3S, seed=74)
    ref = lp.features_batch(w[:12])
    wr, lens = synth.make_ragged(24, 600, 150000, seed=75)
    lens[5] = 0
    good = np.nonzero(lens > 0)[0]
    refr = lp.features_batch(wr[good[:10]], lens[good[:10]])
    try:
        ex.set_pipeline(mode)
        dbg = {}
        got = ex.extract(dev(w), debug=dbg).cpu().numpy()
        assert_parity(got[:12], ref)
        assert (dbg["clip_info"].cpu().numpy()[:, 6] == 0).all()
        assert np.array_equal(ex.extract(dev(w)).cpu().numpy(), got)        # non-debug instantiation, bit for bit
        gr = ex.extract(dev(wr), dev(lens)).cpu().numpy()
        assert np.isnan(gr[5]).all()
        assert_parity(gr[good[:10]], refr)
    finally:
        ex.set_pipeline("auto")


def test_tcgen05_chroma_agrees_with_mma_sync_on_a_large_batch(ex):
    """Mode 4 = the fused kernel with phase 3b on tcgen05 (UMMA, accumulator in tensor memory, operands bulk-copied as
    shared-memory images, remainder frames on the FP32 pipes): everything but chroma bit for bit, chroma to 2e-6; zero-tail
    clips (partial tiles), all-zero clips (no tile at all) and clips of 1 .. 300 frames included."""
    w = synth.make_batch(1200, N3S, seed=77)
    w[5] = 0.0
    wd = dev(w)
    wr, lens = synth.make_ragged(400, 300, 153600, seed=78)
    lens[:8] = [1, 511, 512, 4096 * 16 - 1, 65536, 65536 + 511, 153600, 16384]
    wrd, ld = dev(wr), dev(lens)
    try:
        ex.set_pipeline("fused")
        a, ar = ex.extract(wd), ex.extract(wrd, ld)
        ex.set_pipeline("fused_umma")
        b, br = ex.extract(wd), ex.extract(wrd, ld)
        b2 = ex.extract(wd)
    finally:
        ex.set_pipeline("auto")
    assert torch.equal(b, b2)
    for x, y in ((a, b), (ar, br)):
        assert torch.equal(x[:, :40], y[:, :40]) and torch.equal(x[:, 52:], y[:, 52:])
        assert float((x[:, 40:52] - y[:, 40:52]).abs().max()) < 2e-6
    assert_parity(br[:12].cpu().numpy(), lp.features_batch(wr[:12], lens[:12]))


def test_stream_pipeline_agrees_with_fused_on_a_large_batch(ex):
    """Mode 3 (one persistent 16-warp CTA per SM, warp-granular frame / tail scheduling) runs the arithmetic of the fused kernel:
    MFCC, zcr, centroid, roll-off and rms bit for bit, chroma to 2e-6 (its MMA accumulates the 1 024 bins in a different order);
    the rows do not depend on which warp ran which frame."""
    w = dev(synth.make_batch(1500, N3S, seed=76))
    try:
        ex.set_pipeline("fused")
        a = ex.extract(w)
        ex.set_pipeline("stream")
        b = ex.extract(w)
        c = ex.extract(w)
    finally:
        ex.set_pipeline("auto")
    assert torch.equal(b, c)
    assert torch.equal(a[:, :40], b[:, :40]) and torch.equal(a[:, 52:], b[:, 52:])
    assert float((a[:, 40:52] - b[:, 40:52]).abs().max()) < 2e-6


def test_config5_variable_length_extremes(ex):
    """BASELINE configs[4] shape: 0.5 s .. 60 s, padded rows + lengths (T = 22 .. 2584)."""
    lens = np.array([11025, 11026, 511, 512, 513, 2047, 2048, 2049, 66150, 123457, 1323000], dtype=np.int32)
    rng = np.random.default_rng(13)
    w = np.full((len(lens), 1323000), 3.0, dtype=np.float32)
    for i, n in enumerate(lens):
        w[i, :n] = synth.make_clip(synth.KINDS[i % 4], int(n), rng)
    got = ex.extract(dev(w), dev(lens)).cpu().numpy()
    ref = lp.features_batch(w, lens)
    assert_parity(got, ref)


def test_single_sample_and_bad_lengths(ex):
    w = np.full((3, 4096), 0.5, dtype=np.float32)
    lens = np.array([1, 0, -5], dtype=np.int32)
    got = ex.extract(dev(w), dev(lens)).cpu().numpy()
    assert np.isfinite(got[0]).all()                                # T = 1 frame
    assert_parity(got[:1], lp.features_batch(w[:1], lens[:1]))
    assert np.isnan(got[1:]).all()                                  # length <= 0 -> NaN row (host wrapper raises)
    # a device-side length beyond the row (more frames than the scratch slice holds) is a NaN row as well, in either
    # pipeline, and does not disturb its neighbours
    w = synth.make_batch(6, 4096, seed=5)
    lens = np.array([4096, 4097, 3000, 2**31 - 1, 4096, 100000], dtype=np.int32)
    good = [0, 2, 4]
    try:
        for mode in (1, 2):
            assert ex.lib.sfx_set_pipeline(mode) == 0
            got = ex.extract(dev(w), dev(lens)).cpu().numpy()
            assert np.isnan(got[[1, 3, 5]]).all()
            assert_parity(got[good], lp.features_batch(w[good], lens[good]))
    finally:
        ex.lib.sfx_set_pipeline(0)


def test_unaligned_rows_and_strided_batch(ex):
    base = synth.make_batch(6, N3S + 1, seed=21)
    d = dev(base)
    odd = ex.extract(d, n_samples=N3S).cpu().numpy()                # odd row stride: scalar-load path
    assert_parity(odd, lp.features_batch(base[:, :N3S]))
    shifted = d[:, 1:]                                              # rows start 4 bytes off 8-byte alignment
    got = ex.extract(shifted).cpu().numpy()
    assert_parity(got, lp.features_batch(base[:, 1:]))


@pytest.mark.parametrize("n_mfcc", [13, 20, 128])
def test_other_n_mfcc(ex, n_mfcc):
    w = synth.make_batch(4, N3S, seed=31)
    got = ex.extract(dev(w), n_mfcc=n_mfcc).cpu().numpy()
    assert got.shape == (4, n_mfcc + 16)
    ref = np.stack([np.concatenate([lp.extract_mfcc(x, 22050, n_mfcc), lp.extract_chroma(x, 22050),
                                    lp.extract_spectral_features(x, 22050)]) for x in w])
    assert_parity(got, ref, n_mfcc=n_mfcc)


@pytest.mark.parametrize("sr", [8000, 16000, 44100, 48000])
def test_other_sample_rate(ex, sr):
    """Other table sets: 8 kHz has the widest piptrack range (bins 39..1023, 31 rows of 32) and the fewest mel slots per lane,
    48 kHz the narrowest range and the most slots (mel_ps = 25)."""
    from sfx_b200 import get_extractor
    exs = get_extractor(torch.device("cuda", 0), sr=sr)
    w = synth.make_batch(4, 3 * sr, seed=41)
    try:
        for mode in (1, 2):                                   # fused and split pipelines
            assert exs.lib.sfx_set_pipeline(mode) == 0
            got = exs.extract(dev(w)).cpu().numpy()
            ref = np.stack([lp.features_from_audio(x, sr=sr) for x in w]) if mode == 1 else ref
            assert_parity(got, ref)
    finally:
        exs.lib.sfx_set_pipeline(0)


def test_host_path_matches_device_path_and_chunks(ex):
    w = synth.make_batch(37, N3S, seed=51)
    d = ex.extract(dev(w)).cpu().numpy()
    assert np.array_equal(ex.extract_host(w), d)
    assert np.array_equal(ex.extract_host(w, chunk_clips=8), d)     # 5 chunks over two streams
    pinned = torch.from_numpy(w).pin_memory()
    out = torch.empty((37, 56), dtype=torch.float32).pin_memory()
    ex.extract_host(pinned.numpy(), out=out.numpy(), chunk_clips=16)
    assert np.array_equal(out.numpy(), d)
    wr, lens = synth.make_ragged(9, 11025, 150000, seed=52)
    dr = ex.extract(dev(wr), dev(lens)).cpu().numpy()
    assert np.array_equal(ex.extract_host(wr, lens, chunk_clips=4), dr)


def test_pcm16_host_path_is_the_float_path_on_dequantised_samples(ex):
    """16-bit PCM rows (as in a WAV file at the extractor's rate) are dequantised on the device exactly as soundfile does
    for librosa.load (x / 32768): the features equal, bit for bit, those of the float path fed the dequantised samples."""
    w = synth.make_batch(21, N3S, seed=53)
    pcm = np.clip(np.round(w * 32768.0), -32768, 32767).astype(np.int16)
    deq = pcm.astype(np.float32) / np.float32(32768.0)
    d = ex.extract(dev(deq)).cpu().numpy()
    assert np.array_equal(ex.extract_host(pcm), d)
    assert np.array_equal(ex.extract_host(pcm, chunk_clips=4), d)
    pinned = torch.from_numpy(pcm).pin_memory()
    assert np.array_equal(ex.extract_host(pinned.numpy(), chunk_clips=8), d)
    wr, lens = synth.make_ragged(7, 11025, 90001, seed=54)          # odd maximum length: padded device rows
    pr = np.clip(np.round(wr * 32768.0), -32768, 32767).astype(np.int16)
    dr = ex.extract(dev(pr.astype(np.float32) / np.float32(32768.0)), dev(lens)).cpu().numpy()
    assert np.array_equal(ex.extract_host(pr, lens, chunk_clips=3), dr)
    ok, report = synth.compare(d, lp.features_batch(deq))
    assert ok, "\n" + report


def test_dropin_module_functions(ex):
    """The reference's tests/test_preprocessing.py:30-67 against the drop-in module, plus values vs the oracle."""
    from sfx_b200._config import Config
    from preprocessing.audio_preprocessing import (extract_chroma, extract_mfcc, extract_spectral_features,
                                                   extract_features_batch)
    audio = np.random.default_rng(0).standard_normal(Config.SAMPLE_RATE * Config.AUDIO_DURATION)   # float64
    mfcc = extract_mfcc(audio, Config.SAMPLE_RATE)
    assert mfcc.shape == (Config.N_MFCC,) and np.all(np.isfinite(mfcc))
    chroma = extract_chroma(audio, Config.SAMPLE_RATE)
    assert chroma.shape == (12,) and np.all(np.isfinite(chroma))
    spectral = extract_spectral_features(audio, Config.SAMPLE_RATE)
    assert spectral.shape == (4,) and spectral.dtype == np.float32 and np.all(np.isfinite(spectral))
    a32 = audio.astype(np.float32)
    got = np.concatenate([extract_mfcc(a32, 22050), extract_chroma(a32, 22050), extract_spectral_features(a32, 22050)])
    assert_parity(got[None], lp.features_from_audio(a32)[None])
    batch = extract_features_batch(np.stack([a32, a32 * 0.5]))
    assert np.array_equal(batch[0], got.astype(np.float32))
    with pytest.raises(ValueError):
        extract_features_batch(np.stack([a32, np.full_like(a32, np.nan)]))


def test_preprocess_audio_file_roundtrip(ex, tmp_path):
    import wave
    from preprocessing.audio_preprocessing import load_audio, preprocess_audio, preprocess_audio_batch
    y = synth.make_clip("harmonic", 50000, np.random.default_rng(3))
    p = os.path.join(tmp_path, "a.wav")
    with wave.open(p, "wb") as wf:
        wf.setnchannels(1); wf.setsampwidth(2); wf.setframerate(22050)
        wf.writeframes((y * 32767).astype("<i2").tobytes())
    feats = preprocess_audio(p)
    assert feats.shape == (56,) and feats.dtype == np.float32
    audio, _ = load_audio(p)
    assert_parity(feats[None], lp.features_from_audio(audio)[None])
    fb, kept = preprocess_audio_batch([p, os.path.join(tmp_path, "missing.wav"), p], on_error="skip")
    assert kept == [0, 2] and np.array_equal(fb[0], feats) and np.array_equal(fb[1], feats)


def test_properties_at_full_batch(ex):
    """Size-independent checks on a batch too large for the oracle: determinism, permutation equivariance,
    duplicate rows, amplitude scaling laws (rms ~ a, zcr/chroma/centroid/rolloff invariant, mfcc0 + 20log10(a)*sqrt(128))."""
    B = 4096
    g = torch.Generator(device="cuda").manual_seed(7)
    w = torch.randn((B, N3S), device="cuda", generator=g) * 0.1
    w[1::2] = w[0:1]                                                 # half the batch duplicates clip 0
    a = ex.extract(w)
    b = ex.extract(w)
    assert torch.equal(a, b)                                         # bitwise deterministic
    assert torch.equal(a[1::2], a[0:1].expand(B // 2, -1))           # identical clips -> identical rows
    perm = torch.randperm(B, device="cuda", generator=g)
    assert torch.equal(ex.extract(w[perm].contiguous()), a[perm])    # clips are independent units
    s = ex.extract(w[:512] * 0.25)
    r = a[:512]
    assert torch.allclose(s[:, 55], 0.25 * r[:, 55], rtol=1e-5)
    assert torch.equal(s[:, 52], r[:, 52])
    assert torch.allclose(s[:, 53:55], r[:, 53:55], rtol=1e-4)
    assert torch.allclose(s[:, 40:52], r[:, 40:52], atol=2e-5)       # 0.25 is a power of two: tuning identical
    shift = 20 * np.log10(0.25) * np.sqrt(128.0)
    assert torch.allclose(s[:, 0], r[:, 0] + shift, rtol=1e-5)
    assert torch.allclose(s[:, 1:40], r[:, 1:40], atol=2e-3)
    ref = lp.features_batch(w[:4].cpu().numpy())                     # spot-check against the oracle
    assert_parity(a[:4].cpu().numpy(), ref)


def test_stream_ordering_and_reentrancy(ex):
    """Two extractions on different torch streams with their own workspaces must not interfere."""
    from sfx_b200.extractor import SpeechFeatureExtractor
    ex2 = SpeechFeatureExtractor(torch.device("cuda", 0))
    w1, w2 = dev(synth.make_batch(300, N3S, seed=61)), dev(synth.make_batch(300, N3S, seed=62))
    ref1, ref2 = ex.extract(w1).clone(), ex.extract(w2).clone()
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    with torch.cuda.stream(s1):
        o1 = ex.extract(w1)
    with torch.cuda.stream(s2):
        o2 = ex2.extract(w2)
    torch.cuda.synchronize()
    assert torch.equal(o1, ref1) and torch.equal(o2, ref2)


def torture_clips(n=N3S):
    rng = np.random.default_rng(77)
    t = np.arange(n) / 22050.0
    clips = {}
    clips["int16_speechlike"] = np.round(synth.make_clip("harmonic", n, rng) * 32767) / 32768.0
    clips["very_quiet_noise"] = 1e-5 * rng.standard_normal(n)
    clips["quiet_tone_in_noise"] = 1e-3 * np.sin(2 * np.pi * 440.0 * t) + 1e-6 * rng.standard_normal(n)
    clips["full_scale_clipped_noise"] = np.clip(3.0 * rng.standard_normal(n), -1, 1)
    imp = np.zeros(n); imp[[5000, 20000, 20001, 47000]] = [1.0, -0.7, 0.7, 0.3]
    clips["impulses"] = imp
    clips["nyquist_alternation"] = 0.5 * np.where(np.arange(n) % 2 == 0, 1.0, -1.0)
    clips["chirp_100_to_10k"] = 0.4 * np.sin(2 * np.pi * (100.0 * t + (9900.0 / (2 * t[-1])) * t * t))
    clips["dc_plus_noise"] = 0.3 + 0.01 * rng.standard_normal(n)
    clips["two_tones_close"] = 0.3 * np.sin(2 * np.pi * 1000.0 * t) + 0.3 * np.sin(2 * np.pi * 1012.0 * t)
    burst = np.zeros(n); burst[30000:36000] = 0.5 * rng.standard_normal(6000)
    clips["burst_in_silence"] = burst
    am = 0.5 * np.sin(2 * np.pi * 220.0 * t) * (np.sin(2 * np.pi * 3.0 * t) > 0)
    clips["gated_tone"] = am
    clips["negative_dc"] = np.full(n, -0.2)
    return list(clips), np.stack([v for v in clips.values()]).astype(np.float32)


def test_torture_signals(ex):
    """Numerical edge cases: quantised, very quiet, clipped, impulsive, Nyquist-rate, gated and DC signals.

    `impulses` has exactly flat frame spectra: librosa's piptrack local-max test is then decided by rounding noise alone,
    so its tuning estimate is ill-conditioned by construction (any other FFT library flips it too).  A tuning flip is
    tolerated for that clip only, and its chroma is then checked against the oracle evaluated at the tuning the GPU chose.
    """
    names, w = torture_clips()
    dbg = {}
    got = ex.extract(dev(w), debug=dbg).cpu().numpy()
    ref = lp.features_batch(w)
    assert np.isfinite(got).all()
    tun = dbg["clip_info"].cpu().numpy()[:, 0]
    ref_tun = np.array([lp.debug_intermediates(x)["tuning"] for x in w])
    flipped = [names[i] for i in np.nonzero(np.abs(tun - ref_tun) > 1e-6)[0]]
    assert set(flipped) <= {"impulses"}, list(zip(names, tun, ref_tun))
    edges = np.linspace(-0.5, 0.5, 101)
    for nm in flipped:
        i = names.index(nm)
        t_gpu = float(edges[int(np.argmin(np.abs(edges - tun[i])))])
        ref[i, 40:52] = np.mean(lp.chroma_stft(w[i], tuning=t_gpu).T, axis=0)
    ok, report = synth.compare(got, ref)
    assert ok, "\n" + report + "\n" + str(names)


def test_fast_peak_arithmetic_equals_reference_forms(ex):
    """Phase 2 evaluates librosa's parabolic shift (float64 quotient) and the tuning-residual bin with fast forms guarded
    by exactness checks; the debug kernel re-evaluates every peak with the reference forms and counts disagreements
    (clip_info[:, 6]).  The count must be zero on every kind of signal, in both pipelines."""
    names, wt = torture_clips()
    w = np.concatenate([synth.make_batch(96, N3S, seed=17), wt], axis=0)
    total = 0
    try:
        for mode in (1, 2):
            assert ex.lib.sfx_set_pipeline(mode) == 0
            dbg = {}
            ex.extract(dev(w), debug=dbg)
            ci = dbg["clip_info"].cpu().numpy()
            assert (ci[:, 6] == 0).all(), (mode, np.nonzero(ci[:, 6])[0], ci[ci[:, 6] != 0][:, [2, 6]])
            total += int(ci[:, 2].sum())
    finally:
        ex.lib.sfx_set_pipeline(0)
    assert total > 500_000          # the check saw a meaningful number of peaks
    # a 60 s noise clip: ~250 k peaks overflow the shared-memory key array and the queue of edge cases, so the global
    # key / bin arrays and the inline reference-form path are exercised as well
    rng = np.random.default_rng(5)
    wl = np.zeros((3, 1_323_000), dtype=np.float32)
    lens = np.array([1_323_000, 400_000, 66_150], dtype=np.int32)
    for i, n in enumerate(lens):
        wl[i, :n] = 0.1 * rng.standard_normal(n).astype(np.float32)
    dbg = {}
    ex.extract(dev(wl), dev(lens), debug=dbg)
    ci = dbg["clip_info"].cpu().numpy()
    assert ci[0, 2] > 100_000 and (ci[:, 6] == 0).all(), ci[:, [2, 6]]


def test_ragged_batch_longest_first_order(ex):
    """A ragged batch with more clips than resident CTAs is processed longest clip first (device-side counting sort of
    the clip queue); every clip's row must be what the same clip gives in a small batch that is processed in order."""
    wr, lens = synth.make_ragged(700, 600, 40000, seed=93)
    lens[::97] = 0                                        # invalid clips keep their NaN rows, wherever they are queued
    wd, ld = dev(wr), dev(lens)
    try:
        assert ex.lib.sfx_set_pipeline(1) == 0
        full = ex.extract(wd, ld).cpu().numpy()
        assert ex.lib.sfx_launches_per_extract() == 2     # order kernel + extractor
        parts = [ex.extract(wd[i:i + 100], ld[i:i + 100]).cpu().numpy() for i in range(0, 700, 100)]
        assert ex.lib.sfx_launches_per_extract() == 1
    finally:
        ex.lib.sfx_set_pipeline(0)
    ref = np.concatenate(parts, axis=0)
    assert np.isnan(full[::97]).all() and np.array_equal(np.isnan(full), np.isnan(ref))
    assert np.array_equal(np.nan_to_num(full), np.nan_to_num(ref))


def test_fused_and_split_pipelines_agree(ex):
    """The persistent fused kernel and the frame-parallel two-kernel pipeline run the same arithmetic; only the order of
    the per-clip float64 sums of the per-frame centroid / roll-off differs."""
    w = dev(synth.make_batch(300, N3S, seed=91))
    wr, lens = synth.make_ragged(40, 600, 90000, seed=92)
    wr, lens = dev(wr), dev(lens)
    res = {}
    try:
        for mode, name in ((1, "fused"), (2, "split")):
            assert ex.lib.sfx_set_pipeline(mode) == 0
            res[name] = (ex.extract(w).cpu().numpy(), ex.extract(wr, lens).cpu().numpy(), ex.lib.sfx_launches_per_extract())
    finally:
        ex.lib.sfx_set_pipeline(0)
    assert res["fused"][2] == 1 and res["split"][2] == 3
    for a, b in zip(res["fused"][:2], res["split"][:2]):
        assert np.array_equal(a[:, :53], b[:, :53])                       # mfcc, chroma, zcr: bit-identical
        np.testing.assert_allclose(a[:, 53:55], b[:, 53:55], rtol=3e-7)    # pooled centroid / roll-off
        assert np.array_equal(a[:, 55], b[:, 55])                         # rms
    small = ex.extract(w[:8]).cpu().numpy()                               # auto mode: B <= 256 takes the split pipeline
    assert ex.lib.sfx_launches_per_extract() == 3
    assert np.array_equal(small, res["split"][0][:8])


def test_bench_mix_host_path_equals_fused_kernel(ex):
    """bench.py's pool (noise / harmonic, full length and zero-tailed) through the fused kernel (4 096 clips, device-resident)
    and through the chunked host path (split pipeline): the same bits.  The pool holds clips whose last non-zero frame is
    fainter than 2^-113 -- their 1/scale is 0 like an all-zero frame's, which once moved the end of the chroma projection
    by a frame in the split pipeline only (bench.py's e2e_f32_bitwise_equal_device_path caught it)."""
    import bench
    pool = bench.synth_pool(4096, N3S, seed=1234, device=torch.device("cuda", 0))
    fused = ex.extract(pool).cpu().numpy()
    assert ex.lib.sfx_launches_per_extract() == 1
    host = ex.extract_host(pool.cpu().numpy())
    assert np.array_equal(fused[:, :53], host[:, :53]) and np.array_equal(fused[:, 55], host[:, 55])
    np.testing.assert_allclose(fused[:, 53:55], host[:, 53:55], rtol=3e-7)


def test_median_select_with_massive_duplicates_and_tiny_peak_lists(ex):
    """The median of the peak magnitudes (librosa.pitch_tuning's threshold) on inputs that drive the CTA-wide select through
    all of its exits: hop-periodic signals make every interior frame identical, so thousands of peaks share a handful of
    magnitudes (the surviving bucket never shrinks below the candidate limit and every differing bit is consumed); clips of
    two to five frames hold fewer peaks than the candidate limit from the start.  Fused, split and stream pipelines (the
    stream tail has a select of its own) against the oracle, fused and split bit for bit."""
    sr, n = 22050, N3S
    t = np.arange(n, dtype=np.float64)
    i512 = np.arange(512, dtype=np.float64)

    def tiled(block):                    # one hop of float32 samples repeated: interior frames are identical bit for bit
        return np.tile(block.astype(np.float32), n // 512 + 1)[:n]

    sig = [tiled(0.5 * np.sin(2 * np.pi * 10 * i512 / 512)),
           tiled(0.3 * np.sin(2 * np.pi * 7 * i512 / 512) + 0.2 * np.sin(2 * np.pi * 23 * i512 / 512 + 1.0)),
           tiled(0.4 * np.sign(np.sin(2 * np.pi * i512 / 64.0 + 0.1))),
           0.25 * np.sin(2 * np.pi * (33 * sr / 512) * t / sr) ** 3]
    w = np.stack(sig).astype(np.float32)
    ref = lp.features_batch(w)
    lens = np.array([600, 1100, 2048, 2600], dtype=np.int32)
    rng = np.random.default_rng(5)
    ws = (0.2 * rng.standard_normal((4, 2600))).astype(np.float32)
    ws[0] += (0.3 * np.sin(2 * np.pi * 440.0 * np.arange(2600) / sr)).astype(np.float32)
    ref_s = np.stack([lp.features_batch(ws[i:i + 1, :lens[i]])[0] for i in range(4)])
    res = {}
    try:
        for mode, name in ((1, "fused"), (2, "split"), (3, "stream")):
            assert ex.lib.sfx_set_pipeline(mode) == 0
            res[name] = (ex.extract(dev(w)).cpu().numpy(), ex.extract(dev(ws), dev(lens)).cpu().numpy())
    finally:
        ex.lib.sfx_set_pipeline(0)
    for name, (a, b) in res.items():
        assert_parity(a, ref)
        assert_parity(b, ref_s)
    for a, b in zip(res["fused"], res["split"]):
        assert np.array_equal(a[:, :53], b[:, :53]) and np.array_equal(a[:, 55], b[:, 55])
