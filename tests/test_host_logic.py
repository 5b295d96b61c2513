"""Host-side logic that needs no GPU: the drop-in module's surface and load_audio, clip sharding, and the
world_size-2 all-gather of the feature cache over gloo."""
import inspect
import os
import wave

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sfx_b200 import shard


def test_dropin_surface_matches_reference():
    import preprocessing.audio_preprocessing as ap
    from sfx_b200._config import Config
    assert (Config.SAMPLE_RATE, Config.AUDIO_DURATION, Config.N_MFCC) == (22050, 3, 40)
    sig = {n: str(inspect.signature(getattr(ap, n))) for n in
           ("load_audio", "extract_mfcc", "extract_chroma", "extract_spectral_features", "preprocess_audio")}
    assert sig["load_audio"] == "(file_path, sr=22050, duration=3)"
    assert sig["extract_mfcc"] == "(audio, sr, n_mfcc=40)"
    assert sig["extract_chroma"] == "(audio, sr)"
    assert sig["extract_spectral_features"] == "(audio, sr)"
    assert sig["preprocess_audio"] == "(file_path)"


def test_invalid_audio_raises_before_any_device_work():
    import preprocessing.audio_preprocessing as ap
    with pytest.raises(ValueError):
        ap.extract_mfcc(np.array([0.0, np.inf], dtype=np.float32), 22050)
    with pytest.raises(ValueError):
        ap.extract_chroma(np.zeros(100, dtype=np.int16), 22050)
    with pytest.raises(ValueError):
        ap.extract_spectral_features(np.zeros(0, dtype=np.float32), 22050)


def _write_wav(path, data, rate, width=2):
    with wave.open(path, "wb") as w:
        w.setnchannels(data.shape[1])
        w.setsampwidth(width)
        w.setframerate(rate)
        w.writeframes((data * 32767).astype("<i2").tobytes())


def test_load_audio_pad_trim_and_mono(tmp_path):
    import preprocessing.audio_preprocessing as ap
    rng = np.random.default_rng(0)
    short = rng.uniform(-0.5, 0.5, size=(22050, 2))
    p = os.path.join(tmp_path, "short.wav")
    _write_wav(p, short, 22050)
    audio, sr = ap.load_audio(p)
    assert sr == 22050 and audio.shape == (66150,) and audio.dtype == np.float32
    q = (short * 32767).astype("<i2").astype(np.float32) / 32768.0
    np.testing.assert_allclose(audio[:22050], q.mean(axis=1), atol=1e-7)
    assert not audio[22050:].any()                                    # zero right-pad (reference :15-16)
    long = rng.uniform(-0.5, 0.5, size=(4 * 22050, 1))
    p2 = os.path.join(tmp_path, "long.wav")
    _write_wav(p2, long, 22050)
    audio2, _ = ap.load_audio(p2)
    assert audio2.shape == (66150,)                                   # trimmed (reference :17-18)
    p3 = os.path.join(tmp_path, "rate.wav")
    _write_wav(p3, rng.uniform(-0.5, 0.5, size=(48000, 1)), 48000)
    audio3, sr3 = ap.load_audio(p3)
    assert sr3 == 22050 and audio3.shape == (66150,) and not audio3[22050 + 64:].any()
    # not a WAVE file (an mp3 / ogg upload, reference config.py:49): decoded through soundfile / audioread when importable,
    # else a ValueError subclass that names the missing decoder
    open(os.path.join(tmp_path, "bad.wav"), "wb").write(b"not a wav")
    with pytest.raises(ValueError) as ei:
        ap.load_audio(os.path.join(tmp_path, "bad.wav"))
    try:
        import soundfile  # noqa: F401
    except ImportError:
        try:
            import audioread  # noqa: F401
        except ImportError:
            assert isinstance(ei.value, ap.UnsupportedAudioFormat) and "soundfile" in str(ei.value)


def test_raw_pcm16_reader_selects_only_16_bit_mono_or_stereo(tmp_path):
    """preprocess_audio_batch hands raw frames to the device only for 16-bit PCM mono/stereo files; anything else goes
    through load_audio's host decoder."""
    import wave
    from preprocessing import audio_preprocessing as ap
    rng = np.random.default_rng(2)
    x = rng.integers(-30000, 30000, size=(500, 2)).astype("<i2")
    cases = {"m16": (1, 2, x[:, 0].tobytes()), "s16": (2, 2, x.tobytes()), "m8": (1, 1, (x[:, 0] // 256 + 128).astype("u1").tobytes())}
    for name, (ch, width, payload) in cases.items():
        with wave.open(str(tmp_path / f"{name}.wav"), "wb") as wf:
            wf.setnchannels(ch); wf.setsampwidth(width); wf.setframerate(24414)
            wf.writeframes(payload)
    raw, ch, rate = ap._read_wav_pcm16(str(tmp_path / "m16.wav"))
    assert ch == 1 and rate == 24414 and np.array_equal(raw, x[:, 0])
    raw, ch, rate = ap._read_wav_pcm16(str(tmp_path / "s16.wav"))
    assert ch == 2 and np.array_equal(raw.reshape(-1, 2), x)
    assert ap._read_wav_pcm16(str(tmp_path / "m8.wav")) is None
    (tmp_path / "junk.wav").write_bytes(b"not a wav")
    with pytest.raises(ValueError):
        ap._read_wav_pcm16(str(tmp_path / "junk.wav"))
    # the host decoder agrees with the raw reader on the 16-bit files (x / 32768, channel mean)
    a, sr = ap.load_audio(str(tmp_path / "s16.wav"), sr=24414, duration=1)
    ref = (x.astype(np.float32) / np.float32(32768.0)).mean(axis=1, dtype=np.float32)
    assert sr == 24414 and np.array_equal(a[:500], ref) and not a[500:].any()


def test_shard_ranges_cover_exactly():
    for n, w in ((1_000_000, 8), (1440, 4), (7, 8), (0, 2), (64, 1)):
        ranges = [shard.shard_range(n, w, r) for r in range(w)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        assert max(hi - lo for lo, hi in ranges) == shard.shard_size(n, w)


def test_partition_by_samples_balances():
    rng = np.random.default_rng(0)
    lengths = np.exp(rng.uniform(np.log(11025), np.log(1323000), size=4096)).astype(np.int64)
    parts = shard.partition_by_samples(lengths, 8)
    assert parts[0][0] == 0 and parts[-1][1] == 4096
    loads = np.array([lengths[lo:hi].sum() for lo, hi in parts])
    assert loads.max() / loads.mean() < 1.05


def _worker(rank, world, port, n_total, ragged):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        table = torch.arange(n_total * 5, dtype=torch.float32).reshape(n_total, 5)

        def fake_extract(w, lengths):                                  # stands in for the CUDA extractor
            return torch.cat([w * 2.0, w.sum(dim=1, keepdim=True)], dim=1)[:, :6]

        expect = fake_extract(table, None)
        if ragged:
            lens = np.arange(1, n_total + 1)
            ranges = shard.partition_by_samples(lens, world)
            lo, hi = ranges[rank]
            full = shard.gather_ragged_feature_cache(fake_extract(table[lo:hi], None), ranges)
        else:
            full = shard.extract_sharded(fake_extract, table)
            # the same through the asynchronous form with a caller-owned cache buffer
            lo, hi = shard.shard_range(n_total, world, rank)
            buf = torch.empty((shard.shard_size(n_total, world) * world, 6))
            full2, work = shard.gather_feature_cache(fake_extract(table[lo:hi], None), n_total, out=buf, async_op=True)
            work.wait()
            assert torch.equal(full2, expect) and full2.data_ptr() == buf.data_ptr()
        assert full.shape == expect.shape and torch.equal(full, expect)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total,ragged", [(10, False), (7, False), (9, True)])
def test_feature_cache_allgather_gloo_world2(n_total, ragged):
    port = 29500 + (os.getpid() + n_total) % 2000
    mp.spawn(_worker, args=(2, port, n_total, ragged), nprocs=2, join=True)


REFERENCE = "/root/reference"
PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "multimodal-emotion-classification_b200")


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="the reference checkout exists in the build container only")
def test_documented_dropin_pythonpath_shadows_preprocessing_only():
    """PYTHONPATH=multimodal-emotion-classification_b200:<reference>.  Only `preprocessing` may come
    from this package; `config`, `inference` and `model_training` must resolve to the reference's own modules (its
    SpeechInference reads Config.SPEECH_MODEL_PATH etc. inside try/except and would silently fall back otherwise)."""
    import subprocess
    import sys
    code = (
        "import config, inference.speech_inference as si, preprocessing.audio_preprocessing as ap\n"
        "import model_training, sfx_b200._config as c\n"
        "print(config.__file__); print(si.__file__); print(ap.__file__); print(model_training.__path__[0] if hasattr(model_training, '__path__') else model_training.__file__)\n"
        "assert hasattr(config.Config, 'SPEECH_MODEL_PATH') and hasattr(config.Config, 'SECRET_KEY')\n"
        "assert c.Config is config.Config\n"
        "assert hasattr(si, 'SpeechInference') and si.preprocess_audio is ap.preprocess_audio\n"
        "assert ap.load_audio.__defaults__ == (config.Config.SAMPLE_RATE, config.Config.AUDIO_DURATION)\n"
        "s = si.SpeechInference(); assert s.emotions == config.Config.EMOTIONS\n"
    )
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([PKG, REFERENCE]))
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env, cwd="/tmp")
    assert res.returncode == 0, res.stderr[-2000:]
    cfg, si, ap, mt = res.stdout.strip().splitlines()[-4:]
    assert cfg.startswith(REFERENCE) and si.startswith(REFERENCE) and mt.startswith(REFERENCE)
    assert ap.startswith(PKG)
