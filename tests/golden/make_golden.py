"""Regenerates tests/golden/*.npz from the CPU oracle (oracle/librosa_port.py).

The reference's own tests hold no golden vectors for this path and librosa is not installable here, so these
fixtures pin the ORACLE (regression) rather than librosa itself -- "parity unpinned", see DESIGN.md.
Inputs are regenerated from tests/synth.py seeds; only outputs (+ an input checksum) are stored.
    python tests/golden/make_golden.py
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import synth
from oracle import librosa_port as lp


def main():
    # config 1 (BASELINE.json configs[0]): 64 synthetic 3 s clips, seed 0
    w = synth.make_batch(64, 66150, seed=0)
    feats = lp.features_batch(w)
    tun = np.array([lp.debug_intermediates(x)["tuning"] for x in w])
    np.savez_compressed(os.path.join(HERE, "config1_seed0.npz"), features=feats, tuning=tun,
                        crc=np.uint32(zlib.crc32(w.tobytes())))
    # ragged mini-batch (config 5 shape, small): lengths 0.5 s .. 9 s
    wr, lens = synth.make_ragged(10, 11025, 200000, seed=3)
    np.savez_compressed(os.path.join(HERE, "ragged_seed3.npz"), features=lp.features_batch(wr, lens), lengths=lens,
                        crc=np.uint32(zlib.crc32(wr.tobytes())))
    # edge cases
    rng = np.random.default_rng(5)
    e = np.stack([synth.make_clip(k, 66150, rng) for k in ("zero", "dc", "square")])
    np.savez_compressed(os.path.join(HERE, "edge_cases.npz"), features=lp.features_batch(e),
                        crc=np.uint32(zlib.crc32(e.tobytes())))
    print("golden written")


if __name__ == "__main__":
    main()
