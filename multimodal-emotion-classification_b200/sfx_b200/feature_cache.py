"""
Batched feature-cache builder -- the B200 counterpart of ``load_dataset`` in the reference's
model_training/train_speech_model.py:113-160 (scope row f2).

The reference walks the file list serially and calls preprocess_audio once per file (:121-125).  Here every file is
decoded on the host, the whole batch is extracted in one device pass (chunk-pipelined H2D || kernel || D2H), and with
``torch.distributed`` initialised the file list is clip-sharded over the ranks and the [N, 56] matrix is all-gathered
(NCCL) into every rank's feature cache.  Same signature, prints, label rules and return values (X float32 [N,56],
y one-hot float32 [N,7]); the DNN training that follows in the reference (:169-277) is unchanged and out of scope.

Reference quirk kept on purpose: a file whose features were extracted but whose label cannot be derived
(label_from='name' without a matching key) stays in X while its label is skipped (:125 appends before :127-141 can
raise), and the later ``zip`` (:149) truncates -- identical inputs give identical (X, y).
"""
import glob
import os
import sys
from typing import Dict, List

import numpy as np

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG_ROOT not in sys.path:
    sys.path.insert(0, _PKG_ROOT)

from ._config import Config  # noqa: E402
from preprocessing import audio_preprocessing as _ap  # noqa: E402


def one_hot(labels: List[int], num_classes: int) -> np.ndarray:
    """reference :106-110"""
    y = np.zeros((len(labels), num_classes), dtype=np.float32)
    for i, idx in enumerate(labels):
        y[i, idx] = 1.0
    return y


def _label_of(fp: str, label_from: str, name_map):
    """reference :127-141; raises ValueError exactly where the reference does."""
    if label_from == 'parent':
        return os.path.basename(os.path.dirname(fp)).lower()
    if label_from == 'name':
        base = os.path.basename(fp).lower()
        if name_map:
            for key, val in name_map.items():
                if key.lower() in base:
                    return val
        raise ValueError(f"Could not map filename to label: {base}")
    raise ValueError('label_from must be "parent" or "name"')


def _extract_files(paths):
    """Features of this rank's files in batched device passes: (float32 [n, 56] with NaN rows for failures, {j: error})."""
    if not len(paths):
        return np.zeros((0, 56), dtype=np.float32), {}
    return _ap.preprocess_audio_batch(paths, on_error="collect")


def load_dataset(data_root: str, pattern: str, label_from: str, name_map: Dict[str, str] = None,
                 cache_path: str = None):
    """Load and preprocess the audio dataset (reference :113-160), batched on the GPU.

    cache_path (additive): if given, X and y are also written to ``cache_path`` (.npz) -- the reference recomputes the
    features on every training run."""
    files = glob.glob(os.path.join(data_root, pattern), recursive=True)
    print(f"Found {len(files)} audio files")

    dist = None
    try:
        import torch.distributed as _dist
        if _dist.is_available() and _dist.is_initialized():
            dist = _dist
    except ImportError:
        pass
    world, rank = (dist.get_world_size(), dist.get_rank()) if dist else (1, 0)
    from sfx_b200.shard import shard_range
    lo, hi = shard_range(len(files), world, rank)

    # ---- this rank's files: 16-bit PCM goes to the device as raw frames (dequantise, mono, resample, pad/trim and the
    #      features all happen there), other encodings are decoded on the host; failures are reported and skipped like
    #      the reference's per-file try/except
    ok = np.zeros(len(files), dtype=bool)
    print(f"  Processing {lo}..{hi} of {len(files)}...", end='\r')
    feats_local, local_errors = _extract_files(files[lo:hi])
    errors = {lo + j: e for j, e in local_errors.items()}
    ok[lo:hi] = [j not in local_errors for j in range(hi - lo)]

    # ---- feature cache on every rank
    if dist:
        import torch
        from sfx_b200.shard import gather_feature_cache
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        full = gather_feature_cache(torch.from_numpy(feats_local).to(dev), len(files)).cpu().numpy()
        okt = torch.from_numpy(ok.astype(np.int32)).to(dev)
        dist.all_reduce(okt)
        ok = okt.cpu().numpy() > 0
    else:
        full = feats_local

    X: List[np.ndarray] = []
    y_labels: List[str] = []
    for i, fp in enumerate(files):
        try:
            if not ok[i]:
                raise errors.get(i, RuntimeError("decode failed on another rank"))
            X.append(full[i])
            y_labels.append(_label_of(fp, label_from, name_map))
        except Exception as e:  # noqa: BLE001
            print(f'\nSkip {fp}: {e}')

    print(f"\nSuccessfully processed {len(X)} files")

    # Map labels to indices based on Config.EMOTIONS order
    label_to_idx = {e: i for i, e in enumerate(Config.EMOTIONS)}
    y_idx = [label_to_idx[lbl] for lbl in y_labels if lbl in label_to_idx]
    X = [x for x, lbl in zip(X, y_labels) if lbl in label_to_idx]
    X = np.array(X, dtype=np.float32)
    y = one_hot(y_idx, Config.NUM_EMOTIONS)

    # Print class distribution
    print("\nClass distribution:")
    for emotion in Config.EMOTIONS:
        count = y_labels.count(emotion)
        print(f"  {emotion}: {count} samples")

    if cache_path and rank == 0:
        np.savez_compressed(cache_path, X=X, y=y)
    return X, y
