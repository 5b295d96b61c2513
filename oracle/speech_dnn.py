"""CPU oracle of the speech DNN forward (row f1) -- TEST INFRASTRUCTURE ONLY.

*** PARITY UNPINNED ***  Restates Keras inference of the model built by the reference's
model_training/train_speech_model.py:53-90 (Dense -> BatchNormalization(eps=1e-3) -> ReLU [-> Dropout, inactive at
inference] for widths 512, 512, 256, 128, 64, then Dense(7, softmax)) and the way inference/speech_inference.py:60-105
uses it (StandardScaler.transform, model.predict, arg-max, layers[-3] = the 64-d ReLU output as the fusion feature tap).
TensorFlow/h5py are not installed here and the reference ships no .h5, so weights are synthetic; the arithmetic is
cross-checked against a PyTorch fp32 restatement in tests/test_dnn.py.
"""
from __future__ import annotations

import numpy as np

WIDTHS = (56, 512, 512, 256, 128, 64, 7)       # train_speech_model.py:56-89
BN_EPS = 1e-3                                  # keras.layers.BatchNormalization default epsilon
EMOTIONS = ['happy', 'sad', 'angry', 'fear', 'disgust', 'surprise', 'neutral']   # reference config.py:52


def random_model(seed=0, widths=WIDTHS):
    """Synthetic weights with Keras' initial distributions (Glorot-uniform kernels) and plausible trained BN/scaler state."""
    rng = np.random.default_rng(seed)
    m = {"widths": np.array(widths, dtype=np.int32),
         "scaler_mean": rng.normal(0.0, 50.0, widths[0]).astype(np.float64),          # sklearn stores float64
         "scaler_scale": np.exp(rng.normal(1.0, 1.5, widths[0])).astype(np.float64)}
    for i in range(len(widths) - 1):
        fan_in, fan_out = widths[i], widths[i + 1]
        lim = np.sqrt(6.0 / (fan_in + fan_out))
        m[f"kernel{i}"] = rng.uniform(-lim, lim, (fan_in, fan_out)).astype(np.float32)     # Keras layout [in, out]
        m[f"bias{i}"] = rng.normal(0, 0.05, fan_out).astype(np.float32)
        if i < len(widths) - 2:
            m[f"gamma{i}"] = rng.uniform(0.5, 1.5, fan_out).astype(np.float32)
            m[f"beta{i}"] = rng.normal(0, 0.1, fan_out).astype(np.float32)
            m[f"mean{i}"] = rng.normal(0, 0.3, fan_out).astype(np.float32)
            m[f"var{i}"] = rng.uniform(0.2, 2.0, fan_out).astype(np.float32)
    return m


def scaler_transform(model, X):
    """sklearn StandardScaler.transform on float32 rows (speech_inference.py:66-67): (X - mean_) / scale_ in float32."""
    X = np.array(X, dtype=np.float32, copy=True)
    X -= model["scaler_mean"]            # in-place on float32: each result is rounded to float32
    X /= model["scaler_scale"]
    return X


def forward(model, feats_scaled):
    """Keras model.predict in float32: returns (probs [B,7], tap [B,64] = layers[-3] output)."""
    x = np.asarray(feats_scaled, dtype=np.float32)
    n = len(model["widths"]) - 1
    tap = None
    for i in range(n):
        x = x @ model[f"kernel{i}"] + model[f"bias{i}"]
        if i < n - 1:
            inv = (model[f"gamma{i}"] / np.sqrt(model[f"var{i}"] + np.float32(BN_EPS))).astype(np.float32)
            x = (x - model[f"mean{i}"]) * inv + model[f"beta{i}"]
            x = np.maximum(x, 0.0).astype(np.float32)
            tap = x
    z = x - x.max(axis=1, keepdims=True)
    e = np.exp(z)
    return (e / e.sum(axis=1, keepdims=True)).astype(np.float32), tap


def predict(model, feats):
    """speech_inference.py:60-77 on a batch of raw 56-d feature rows: list of result dicts."""
    probs, _ = forward(model, scaler_transform(model, feats))
    out = []
    for p in probs:
        i = int(np.argmax(p))
        out.append({"emotion": EMOTIONS[i], "confidence": float(p[i]), "all_probabilities": p.tolist()})
    return out
