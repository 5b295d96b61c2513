"""Scope row f2: the batched load_dataset.  CPU: file walking / label rules / skip semantics against a transliteration
of the reference loop (train_speech_model.py:113-160) with a stub extractor.  GPU: real WAV files end to end."""
import glob
import os
import wave

import numpy as np
import pytest

import synth
from oracle import librosa_port as lp

EMO = ['happy', 'sad', 'angry', 'fear', 'disgust', 'surprise', 'neutral']


def write_wav(path, y, rate=22050):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with wave.open(path, "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(rate)
        w.writeframes((np.clip(y, -1, 1) * 32767).astype("<i2").tobytes())


def make_tree(root, n_per=2, seed=0):
    rng = np.random.default_rng(seed)
    for e in ("happy", "sad", "neutral", "bored"):                      # "bored" is not in Config.EMOTIONS
        for i in range(n_per):
            write_wav(os.path.join(root, e, f"{e}_{i}.wav"), synth.make_clip(synth.KINDS[i % 4], 30000 + 5000 * i, rng))
    with open(os.path.join(root, "sad", "broken.wav"), "wb") as fh:
        fh.write(b"not a wav file")


def reference_loop(files, feature_fn, label_from, name_map):
    """Transliteration of the reference's serial loop (same list/zip semantics), feature_fn = preprocess_audio."""
    X, y_labels = [], []
    for fp in files:
        try:
            X.append(feature_fn(fp))
            if label_from == 'parent':
                lbl = os.path.basename(os.path.dirname(fp)).lower()
            else:
                base = os.path.basename(fp).lower()
                lbl = None
                for key, val in (name_map or {}).items():
                    if key.lower() in base:
                        lbl = val
                        break
                if lbl is None:
                    raise ValueError("no label")
            y_labels.append(lbl)
        except Exception:
            pass
    idx = {e: i for i, e in enumerate(EMO)}
    y_idx = [idx[l] for l in y_labels if l in idx]
    X = np.array([x for x, l in zip(X, y_labels) if l in idx], dtype=np.float32)
    y = np.zeros((len(y_idx), 7), dtype=np.float32)
    y[np.arange(len(y_idx)), y_idx] = 1.0
    return X, y


def stub_features(waves):
    w = np.asarray(waves, dtype=np.float32)
    return np.stack([np.concatenate([[x.sum(), np.abs(x).max()], np.arange(54)]) for x in w]).astype(np.float32)


@pytest.mark.parametrize("label_from,name_map", [("parent", None), ("name", {"happy": "happy", "sad": "sad"})])
def test_load_dataset_semantics_match_reference_loop(tmp_path, monkeypatch, capsys, label_from, name_map):
    from sfx_b200 import feature_cache as tsm
    import preprocessing.audio_preprocessing as ap
    make_tree(str(tmp_path))

    def stub_extract_files(paths):                      # stands in for the device passes of preprocess_audio_batch
        feats = np.full((len(paths), 56), np.nan, dtype=np.float32)
        errors = {}
        for j, fp in enumerate(paths):
            try:
                audio, _ = ap.load_audio(fp)
                feats[j] = stub_features(audio[None])[0]
            except Exception as e:  # noqa: BLE001
                errors[j] = e
        return feats, errors

    monkeypatch.setattr(tsm, "_extract_files", stub_extract_files)
    X, y = tsm.load_dataset(str(tmp_path), "**/*.wav", label_from, name_map, cache_path=os.path.join(tmp_path, "cache.npz"))
    files = glob.glob(os.path.join(str(tmp_path), "**/*.wav"), recursive=True)

    def feat(fp):
        audio, _ = ap.load_audio(fp)
        return stub_features(audio[None])[0]

    Xr, yr = reference_loop(files, feat, label_from, name_map)
    assert X.dtype == np.float32 and y.dtype == np.float32
    assert np.array_equal(X, Xr) and np.array_equal(y, yr)
    out = capsys.readouterr().out
    assert "Found 9 audio files" in out and "Skip" in out and "Class distribution" in out
    c = np.load(os.path.join(tmp_path, "cache.npz"))
    assert np.array_equal(c["X"], X) and np.array_equal(c["y"], y)


@pytest.mark.gpu
def test_load_dataset_end_to_end_on_gpu(tmp_path):
    from sfx_b200 import feature_cache as tsm
    import preprocessing.audio_preprocessing as ap
    make_tree(str(tmp_path), n_per=3, seed=4)
    X, y = tsm.load_dataset(str(tmp_path), "**/*.wav", "parent")
    files = glob.glob(os.path.join(str(tmp_path), "**/*.wav"), recursive=True)

    def feat(fp):
        audio, _ = ap.load_audio(fp)
        return lp.features_from_audio(audio)

    Xr, yr = reference_loop(files, feat, "parent", None)
    assert X.shape == Xr.shape == (9, 56) and np.array_equal(y, yr)
    ok, rep = synth.compare(X, Xr)
    assert ok, rep
