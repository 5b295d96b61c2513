"""Pin the oracle (oracle/librosa_port.py) against librosa itself -- one command for any machine where `import librosa`
works (it does not in the build container: no wheel, no network; reference requirements.txt:11 pins librosa==0.10.0).

    python tools/pin_against_librosa.py [--reference /path/to/reference/repo] [--out profiles/librosa_pin.json]

Runs the reference's own extraction on the committed golden inputs (tests/golden: config 1's 64 clips, the ragged batch, the
edge cases; regenerated from the seeded generator and checked against the fixtures' CRC) and reports, per feature group, the
maximum error of the oracle and of the committed fixtures against librosa, in units of the test tolerance and as pure relative
error, plus the number of tuning (chroma filter bank) disagreements.  With --reference the five functions are imported from the
reference's preprocessing/audio_preprocessing.py (:22-37); without it the same librosa calls are made directly.
Exit code 0 = oracle within tolerance of librosa everywhere, 1 = not, 2 = librosa not importable.
"""
import argparse
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default=None, help="checkout of RachaCodez/multimodal-emotion-classification")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "librosa_pin.json"))
    a = ap.parse_args()
    try:
        import librosa
    except Exception as e:      # noqa: BLE001
        print(f"librosa is not importable here ({e}); nothing pinned")
        return 2
    import synth
    from oracle import librosa_port as lp
    if a.reference:
        sys.path.insert(0, a.reference)
        from preprocessing.audio_preprocessing import extract_chroma, extract_mfcc, extract_spectral_features
    else:
        def extract_mfcc(audio, sr, n_mfcc=40):                 # reference preprocessing/audio_preprocessing.py:22-24
            return np.mean(librosa.feature.mfcc(y=audio, sr=sr, n_mfcc=n_mfcc).T, axis=0)

        def extract_chroma(audio, sr):                           # :27-29
            return np.mean(librosa.feature.chroma_stft(y=audio, sr=sr).T, axis=0)

        def extract_spectral_features(audio, sr):                # :32-37
            zcr = np.mean(librosa.feature.zero_crossing_rate(audio))
            centroid = np.mean(librosa.feature.spectral_centroid(y=audio, sr=sr))
            rolloff = np.mean(librosa.feature.spectral_rolloff(y=audio, sr=sr))
            rms = np.mean(librosa.feature.rms(y=audio))
            return np.array([zcr, centroid, rolloff, rms], dtype=np.float32)

    def reference_rows(waves, lengths=None):
        out = np.empty((len(waves), 56), dtype=np.float32)
        for i, w in enumerate(waves):
            y = np.ascontiguousarray(w if lengths is None else w[:int(lengths[i])])
            out[i] = np.concatenate([extract_mfcc(y, 22050), extract_chroma(y, 22050), extract_spectral_features(y, 22050)])
        return out

    gold = os.path.join(ROOT, "tests", "golden")
    sets = {}
    g1 = np.load(os.path.join(gold, "config1_seed0.npz"))
    w1 = synth.make_batch(64, 66150, seed=0)
    assert np.uint32(zlib.crc32(w1.tobytes())) == g1["crc"], "synthetic generator drifted from the fixtures"
    sets["config1_seed0"] = (w1, None, g1["features"])
    gr = np.load(os.path.join(gold, "ragged_seed3.npz"))
    wr, lens = synth.make_ragged(10, 11025, 200000, seed=3)
    sets["ragged_seed3"] = (wr, lens, gr["features"])
    ge = np.load(os.path.join(gold, "edge_cases.npz"))
    rng = np.random.default_rng(5)
    we = np.stack([synth.make_clip(k, 66150, rng) for k in ("zero", "dc", "square")])
    sets["edge_cases"] = (we, None, ge["features"])

    report = {"librosa": librosa.__version__, "numpy": np.__version__, "reference_module": bool(a.reference), "sets": {}}
    all_ok = True
    for name, (w, ln, fixture) in sets.items():
        ref = reference_rows(w, ln)
        ora = lp.features_batch(w, ln)
        ok_o, rep_o = synth.compare(ora, ref)
        ok_f, rep_f = synth.compare(fixture, ref)
        tun_ref = [float(librosa.estimate_tuning(y=np.ascontiguousarray(x if ln is None else x[:int(ln[i])]), sr=22050, bins_per_octave=12))
                   if np.any(x) else 0.0 for i, x in enumerate(w)]
        tun_ora = [lp.debug_intermediates(np.ascontiguousarray(x if ln is None else x[:int(ln[i])]))["tuning"] for i, x in enumerate(w)]
        rel = np.abs(ora.astype(np.float64) - ref) / np.maximum(np.abs(ref.astype(np.float64)), 1e-30)
        report["sets"][name] = {"clips": int(len(w)), "oracle_within_tolerance": ok_o, "fixture_within_tolerance": ok_f,
                                "oracle_vs_librosa": rep_o.splitlines(), "fixture_vs_librosa": rep_f.splitlines(),
                                "tuning_mismatches": int(sum(abs(x - y) > 1e-9 for x, y in zip(tun_ref, tun_ora))),
                                "oracle_bitwise_equal": bool(np.array_equal(ora, ref)),
                                "pure_relative_error_max": float(rel.max()), "pure_relative_error_p99": float(np.quantile(rel, 0.99))}
        all_ok &= ok_o and ok_f
        print(name, "oracle ok" if ok_o else "ORACLE DIFFERS", "| fixture ok" if ok_f else "| FIXTURE DIFFERS")
        print(rep_o)
    report["pinned"] = bool(all_ok)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(report, open(a.out, "w"), indent=1)
    print("written", a.out)
    return 0 if all_ok else 1


if __name__ == "__main__":
    sys.exit(main())
