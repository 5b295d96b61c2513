"""Reference-side exporter: the reference's trained artefacts -> the .npz that sfx_b200.inference.BatchedSpeechInference loads.

Run it where the reference runs (TensorFlow + joblib installed, reference requirements.txt); this build's container has
neither, which is why the conversion lives on that side of the boundary:

    python tools/export_weights.py --h5 models/speech_model.h5 --scaler models/speech_scaler.pkl --out speech_model.npz

Reads what inference/speech_inference.py:17-34 loads (tf.keras.models.load_model(Config.SPEECH_MODEL_PATH),
joblib.load(Config.SPEECH_SCALER_PATH)) and writes, for the architecture of model_training/train_speech_model.py:53-90
(Dense -> BatchNormalization -> ReLU -> Dropout, five times, then Dense softmax):
    widths int32[n+1]; kernel{i} float32[in,out], bias{i} float32[out]   (Keras Dense layout, unchanged)
    gamma{i}, beta{i}, mean{i}, var{i} float32[out]                       (BatchNormalization after hidden Dense i)
    bn_eps float32; scaler_mean, scaler_scale float64[56]                 (sklearn StandardScaler.mean_ / .scale_)
`layers_to_npz` works on any object with a Keras-like `.layers` list (used by tests/test_dnn.py with stand-in layers).
"""
import argparse

import numpy as np


def layers_to_npz(layers, scaler=None) -> dict:
    """Walk a Keras-style layer list: every Dense starts a block, the BatchNormalization that follows it (if any) belongs to
    it; Activation / Dropout / InputLayer carry no weights."""
    out, widths, i = {}, [], -1
    eps = None
    for layer in layers:
        kind = type(layer).__name__
        w = layer.get_weights()
        if kind == "Dense":
            i += 1
            kernel, bias = w if len(w) == 2 else (w[0], np.zeros(w[0].shape[1], np.float32))
            if not widths:
                widths.append(int(kernel.shape[0]))
            if widths[-1] != kernel.shape[0]:
                raise ValueError(f"Dense {i}: input width {kernel.shape[0]} does not follow {widths[-1]}")
            widths.append(int(kernel.shape[1]))
            out[f"kernel{i}"] = np.asarray(kernel, np.float32)
            out[f"bias{i}"] = np.asarray(bias, np.float32)
        elif kind == "BatchNormalization":
            if i < 0 or len(w) != 4:
                raise ValueError("BatchNormalization without a preceding Dense, or without scale/center")
            gamma, beta, mean, var = w
            out[f"gamma{i}"], out[f"beta{i}"] = np.asarray(gamma, np.float32), np.asarray(beta, np.float32)
            out[f"mean{i}"], out[f"var{i}"] = np.asarray(mean, np.float32), np.asarray(var, np.float32)
            e = float(getattr(layer, "epsilon", 1e-3))
            if eps is not None and e != eps:
                raise ValueError("BatchNormalization layers with different epsilon")
            eps = e
        elif w:
            raise ValueError(f"layer type {kind} carries weights and is not part of the speech DNN")
    n = i + 1
    if n < 1:
        raise ValueError("no Dense layer found")
    if any(f"gamma{k}" in out for k in range(n)) and not all(f"gamma{k}" in out for k in range(n - 1)):
        raise ValueError("every hidden Dense must be followed by BatchNormalization (or none)")
    out["widths"] = np.array(widths, np.int32)
    out["bn_eps"] = np.float32(1e-3 if eps is None else eps)
    if scaler is not None:
        out["scaler_mean"] = np.asarray(scaler.mean_, np.float64)
        out["scaler_scale"] = np.asarray(scaler.scale_, np.float64)
        if out["scaler_mean"].shape != (widths[0],):
            raise ValueError("scaler width does not match the model input")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--h5", required=True)
    ap.add_argument("--scaler", default=None)
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    import tensorflow as tf
    model = tf.keras.models.load_model(a.h5)
    scaler = None
    if a.scaler:
        import joblib
        scaler = joblib.load(a.scaler)
    d = layers_to_npz(model.layers, scaler)
    np.savez(a.out, **d)
    print("written", a.out, "widths", d["widths"].tolist(), "scaler" if scaler is not None else "no scaler")


if __name__ == "__main__":
    main()
