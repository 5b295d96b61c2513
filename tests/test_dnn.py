"""Scope row f1: scaler + speech DNN forward.  CPU: the numpy oracle against an independent PyTorch restatement.
GPU (-m gpu): sfx_dnn_forward through the C ABI against the oracle, and config 3 (TESS-shaped clips -> features -> DNN
with device-resident features)."""
import os

import numpy as np
import pytest
import torch

import synth
from oracle import librosa_port as lp
from oracle import speech_dnn as od


def torch_reference(model, feats):
    x = torch.from_numpy(np.asarray(feats, dtype=np.float32)).double()
    x = (x - torch.from_numpy(model["scaler_mean"])) / torch.from_numpy(model["scaler_scale"])
    n = len(model["widths"]) - 1
    tap = None
    for i in range(n):
        x = torch.nn.functional.linear(x, torch.from_numpy(model[f"kernel{i}"]).double().T,
                                       torch.from_numpy(model[f"bias{i}"]).double())
        if i < n - 1:
            x = torch.nn.functional.batch_norm(
                x, torch.from_numpy(model[f"mean{i}"]).double(), torch.from_numpy(model[f"var{i}"]).double(),
                torch.from_numpy(model[f"gamma{i}"]).double(), torch.from_numpy(model[f"beta{i}"]).double(),
                training=False, eps=od.BN_EPS)
            x = torch.relu(x)
            tap = x
    return torch.softmax(x, dim=1).numpy(), tap.numpy()


def feature_like(B, seed=0, model_seed=0):
    rng = np.random.default_rng(seed)
    m = od.random_model(model_seed)
    return (m["scaler_mean"] + m["scaler_scale"] * rng.standard_normal((B, 56))).astype(np.float32)


def test_oracle_matches_torch_reference():
    m = od.random_model(1)
    X = feature_like(64, 2, 1)
    probs, tap = od.forward(m, od.scaler_transform(m, X))
    rp, rt = torch_reference(m, X)
    assert probs.shape == (64, 7) and tap.shape == (64, 64)
    assert np.abs(probs - rp).max() < 2e-5 and np.abs(tap - rt).max() < 2e-4 * max(1.0, np.abs(rt).max())
    assert np.allclose(probs.sum(axis=1), 1.0, atol=1e-6)
    res = od.predict(m, X[:3])
    assert set(res[0]) == {"emotion", "confidence", "all_probabilities"} and res[0]["emotion"] in od.EMOTIONS


@pytest.mark.gpu
def test_dnn_forward_matches_oracle():
    from sfx_b200.dnn import SpeechDNN
    m = od.random_model(3)
    dnn = SpeechDNN(m, torch.device("cuda", 0))
    for B in (1, 7, 64, 1000):
        X = feature_like(B, 10 + B, 3)
        probs, tap = dnn.forward(torch.from_numpy(X).cuda())
        rp, rt = od.forward(m, od.scaler_transform(m, X))
        assert np.abs(probs.cpu().numpy() - rp).max() < 2e-5
        assert np.abs(tap.cpu().numpy() - rt).max() < 5e-4 * max(1.0, np.abs(rt).max())
        assert (probs.argmax(dim=1).cpu().numpy() == rp.argmax(axis=1)).mean() > 0.99
    assert "libsfx_b200.so" in open("/proc/self/maps").read()


@pytest.mark.gpu
def test_config3_tess_shaped_features_feed_dnn_on_device():
    """BASELINE configs[2]: 2 800 synthetic ~2 s clips zero-padded to 3 s; features stay on the GPU and feed the DNN."""
    from sfx_b200.inference import BatchedSpeechInference
    rng = np.random.default_rng(33)
    B, n = 2800, 66150
    g = torch.Generator(device="cuda").manual_seed(5)
    w = torch.randn((B, n), device="cuda", generator=g) * 0.1
    cut = torch.from_numpy((44100 * rng.uniform(0.85, 1.15, B)).astype(np.int64)).cuda()
    w.masked_fill_(torch.arange(n, device="cuda")[None, :] >= cut[:, None], 0.0)
    m = od.random_model(4)
    eng = BatchedSpeechInference(m, torch.device("cuda", 0))
    feats = eng.extractor.extract(w)
    # a scaler fitted to these features, as train_speech_model.py:196-198 does
    m["scaler_mean"] = feats.double().mean(dim=0).cpu().numpy()
    m["scaler_scale"] = feats.double().std(dim=0).clamp_min(1e-6).cpu().numpy()
    eng = BatchedSpeechInference(m, torch.device("cuda", 0))
    probs, tap, feats = eng.forward(w)
    assert probs.is_cuda and tap.is_cuda and feats.is_cuda and probs.shape == (B, 7) and tap.shape == (B, 64)
    f = feats.cpu().numpy()
    rp, rt = od.forward(m, od.scaler_transform(m, f))                     # DNN parity on the device features
    assert np.abs(probs.cpu().numpy() - rp).max() < 5e-5
    idx = rng.choice(B, size=6, replace=False)                            # feature parity on a sample of clips
    ok, rep = synth.compare(f[idx], lp.features_batch(w[idx].cpu().numpy()))
    assert ok, rep
    res = eng.predict_batch(w[:4])
    assert [r["emotion"] for r in res] == [od.EMOTIONS[i] for i in rp[:4].argmax(axis=1)]
    t64, p7 = eng.extract_features_batch(w[:4])
    assert t64.shape == (4, 64) and p7.shape == (4, 7)


@pytest.mark.gpu
def test_heuristic_predict_batch_thresholds():
    """speech_inference.py:36-58: rms > 0.06 & centroid > 2000 -> angry; rms < 0.02 & centroid < 1500 -> sad; else neutral."""
    from sfx_b200.inference import BatchedSpeechInference
    n = 66150
    t = np.arange(n) / 22050.0
    loud_bright = (0.3 * np.random.default_rng(0).standard_normal(n)).clip(-1, 1).astype(np.float32)      # rms .3, centroid ~5.5k
    quiet_low = (0.01 * np.sin(2 * np.pi * 200 * t)).astype(np.float32)                                   # rms .007, centroid ~200
    mid = (0.05 * np.sin(2 * np.pi * 300 * t)).astype(np.float32)                                         # rms .035
    eng = BatchedSpeechInference(od.random_model(0), torch.device("cuda", 0))
    res = eng.heuristic_predict_batch(torch.from_numpy(np.stack([loud_bright, quiet_low, mid])).cuda())
    assert [r["emotion"] for r in res] == ["angry", "sad", "neutral"]
    assert all(abs(sum(r["all_probabilities"]) - 1.0) < 1e-9 and r["confidence"] == 0.9 for r in res)
    spec = lp.extract_spectral_features(quiet_low, 22050)
    assert spec[3] < 0.02 and spec[1] < 1500


def test_export_weights_walks_a_keras_style_layer_list(tmp_path):
    """tools/export_weights.py (reference side): Dense / BatchNormalization / Activation / Dropout stand-ins with Keras'
    get_weights() orders -> the dict SpeechDNN takes; round trip through .npz and the oracle forward."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import export_weights as ew
    m = od.random_model(9)

    def layer(kind, weights, **attrs):
        cls = type(kind, (), {"get_weights": lambda self: weights})
        obj = cls()
        for k, v in attrs.items():
            setattr(obj, k, v)
        return obj

    layers = [layer("InputLayer", [])]
    n = len(m["widths"]) - 1
    for i in range(n):
        layers.append(layer("Dense", [m[f"kernel{i}"], m[f"bias{i}"]]))
        if i < n - 1:
            layers.append(layer("BatchNormalization", [m[f"gamma{i}"], m[f"beta{i}"], m[f"mean{i}"], m[f"var{i}"]], epsilon=1e-3))
            layers.append(layer("Activation", []))
            layers.append(layer("Dropout", []))
    scaler = type("StandardScaler", (), {})()
    scaler.mean_, scaler.scale_ = m["scaler_mean"], m["scaler_scale"]
    d = ew.layers_to_npz(layers, scaler)
    path = os.path.join(tmp_path, "speech_model.npz")
    np.savez(path, **d)
    from sfx_b200.inference import load_weights
    back = load_weights(path)
    assert back["widths"].tolist() == list(od.WIDTHS) and float(back["bn_eps"]) == pytest.approx(1e-3)
    for k, v in m.items():
        assert np.array_equal(back[k], v), k
    X = feature_like(5, 1, 2)
    a = od.forward(m, od.scaler_transform(m, X))
    b = od.forward(back, od.scaler_transform(back, X))
    assert np.array_equal(a[0], b[0])
    with pytest.raises(ValueError):
        ew.layers_to_npz(layers[:3] + [layer("Conv1D", [np.zeros(3)])])
