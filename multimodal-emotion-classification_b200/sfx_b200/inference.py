"""
Batched, device-resident speech inference -- the B200 counterpart of the reference's inference/speech_inference.py.

The reference builds features per file with librosa, scales them with a joblib'd StandardScaler and runs a Keras .h5
(speech_inference.py:60-105).  Here a whole batch of clips goes waveform -> 56-d features -> scaler -> DNN on the GPU
(sfx_extract + sfx_dnn_forward) and only the 7 probabilities (and optionally the 64-d fusion tap) come back.
The result dictionaries have the reference's keys.  Weights are a dict of numpy arrays (see sfx_b200/dnn.py) or an .npz written by
tools/export_weights.py, which runs on the reference side (where TensorFlow / joblib exist) and converts the reference's
models/speech_model.h5 + speech_scaler.pkl (speech_inference.py:17-34).
"""
from typing import Dict, List

import numpy as np

from ._config import Config


def load_weights(path: str) -> dict:
    """The .npz written by tools/export_weights.py (from the reference's speech_model.h5 + speech_scaler.pkl)."""
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


class BatchedSpeechInference:
    def __init__(self, weights, device=None, sr: int = Config.SAMPLE_RATE):
        """weights: dict of numpy arrays (sfx_b200/dnn.py) or the path of an .npz from tools/export_weights.py."""
        import torch
        if isinstance(weights, (str, bytes)) or hasattr(weights, "__fspath__"):
            weights = load_weights(weights)
        from . import get_extractor
        from .dnn import SpeechDNN
        self.emotions = Config.EMOTIONS
        dev = torch.device("cuda") if device is None else torch.device(device)
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        self.extractor = get_extractor(self.device, sr)
        self.dnn = SpeechDNN(weights, self.device, bn_eps=float(weights.get("bn_eps", 1e-3)))

    def forward(self, waves, lengths=None, n_samples=None):
        """cuda float32 [B, L] -> (probs [B,7], tap [B,64], features [B,56]), all device tensors, stream-ordered."""
        feats = self.extractor.extract(waves, lengths, n_samples=n_samples)
        probs, tap = self.dnn.forward(feats)
        return probs, tap, feats

    def predict_batch(self, waves, lengths=None) -> List[Dict]:
        """Reference SpeechInference.predict (:60-77) for every clip of the batch."""
        probs, _, _ = self.forward(waves, lengths)
        p = probs.cpu().numpy()
        out = []
        for row in p:
            idx = int(np.argmax(row))
            out.append({'emotion': self.emotions[idx], 'confidence': float(row[idx]), 'all_probabilities': row.tolist()})
        return out

    def heuristic_predict_batch(self, waves, lengths=None) -> List[Dict]:
        """Reference SpeechInference._heuristic_predict (:36-58) for a batch: thresholds on the frame-mean rms and
        spectral centroid (columns 55 and 53 of the feature rows); used by the reference when no .h5 is present."""
        feats = self.extractor.extract(waves, lengths)
        spec = feats[:, Config.N_MFCC + 12:Config.N_MFCC + 16].cpu().numpy()
        out = []
        for zcr, centroid, rolloff, rms in spec:
            if rms > 0.06 and centroid > 2000:
                label = 'angry'
            elif rms < 0.02 and centroid < 1500:
                label = 'sad'
            else:
                label = 'neutral'
            probs = np.ones(len(self.emotions)) * (0.1 / (len(self.emotions) - 1))
            idx = self.emotions.index(label)
            probs[idx] = 0.9
            out.append({'emotion': label, 'confidence': float(probs[idx]), 'all_probabilities': probs.tolist()})
        return out

    def extract_features_batch(self, waves, lengths=None):
        """Reference SpeechInference.extract_features (:79-105): (intermediate [B,64], predictions [B,7]) as numpy."""
        probs, tap, _ = self.forward(waves, lengths)
        return tap.cpu().numpy(), probs.cpu().numpy()
