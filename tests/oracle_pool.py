"""Oracle over many clips on all host cores (test infrastructure): every worker synthesises its own clips from the seeded
generator of tests/synth.py and runs oracle/librosa_port.py on them, so nothing big crosses a process boundary.

    feats, waves = oracle_batch(B, n, seed, ...)        # what synth.make_batch(B, n, seed) would give, and its oracle rows
    feats = oracle_rows(waves, lengths)                 # oracle rows of given clips
"""
import multiprocessing as mp
import os

import numpy as np

import synth
from oracle import librosa_port as lp


def cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def _rows_worker(args):
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    waves, lengths = args
    return lp.features_batch(waves, lengths)


def oracle_rows(waves, lengths=None, procs=None, ctx="fork"):
    """Oracle feature rows of `waves` ([B, n] float32; row i uses waves[i, :lengths[i]]), split over a process pool
    (ctx='spawn' for callers that have initialised CUDA: bench.py)."""
    B = len(waves)
    procs = min(procs or cores(), max(1, B))
    if procs == 1:
        return lp.features_batch(waves, lengths)
    if lengths is None:
        order = np.arange(B)
    else:
        order = np.argsort(-np.asarray(lengths))                         # longest first, dealt round-robin: balanced chunks
    chunks = [order[i::procs] for i in range(procs)]
    jobs = [(np.ascontiguousarray(waves[c] if lengths is None else waves[c][:, :int(np.max(np.asarray(lengths)[c]))]),
             None if lengths is None else np.asarray(lengths)[c]) for c in chunks if len(c)]
    with mp.get_context(ctx).Pool(len(jobs)) as pool:
        parts = pool.map(_rows_worker, jobs)
    out = np.empty((B, parts[0].shape[1]), dtype=np.float32)
    for c, part in zip([c for c in chunks if len(c)], parts):
        out[c] = part
    return out


def _make_clip_i(args):
    kinds, n, seed, i = args
    rng = np.random.default_rng([seed, i])
    return synth.make_clip(kinds[i % len(kinds)], n, rng)


def _batch_worker(args):
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    kinds, n, n_used, seed, idx, want_waves = args
    waves = np.stack([_make_clip_i((kinds, n, seed, i)) for i in idx])
    feats = [lp.features_batch(waves[:, :m]) for m in n_used]
    return idx, (waves if want_waves else None), feats


def indexed_batch(B, n, seed, kinds=synth.KINDS, n_used=None, procs=None, want_waves=True):
    """B clips, clip i drawn from default_rng([seed, i]) (so any subset can be regenerated on its own), and the oracle rows
    of clip[:m] for every m in `n_used` (default: the whole clip).  Returns (waves [B, n] or None, [feats per m])."""
    n_used = [n] if n_used is None else list(n_used)
    procs = min(procs or cores(), B)
    jobs = [(kinds, n, n_used, seed, list(range(p, B, procs)), want_waves) for p in range(procs)]
    with mp.get_context("fork").Pool(procs) as pool:
        res = pool.map(_batch_worker, jobs)
    waves = np.empty((B, n), dtype=np.float32) if want_waves else None
    feats = [np.empty((B, 56), dtype=np.float32) for _ in n_used]
    for idx, w, fs in res:
        if want_waves:
            waves[idx] = w
        for k, f in enumerate(fs):
            feats[k][idx] = f
    return waves, feats
