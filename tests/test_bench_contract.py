"""bench.py's CPU-checkable parts: the reference arm prints one JSON line with the contract's keys on the main arm's
config, and the algorithmic byte / flop figures behind the two rooflines are the ones DESIGN.md states."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    import bench
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == "clips/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    assert d["config"]["workload"] == bench.WORKLOAD and d["config"]["n_samples"] == 66150
    assert d["cpu_baseline"]["kind"] == ("reference" if bench.librosa_available() else "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["gpu_launches"] == 0


def test_algorithmic_bytes_and_flops():
    import bench
    assert bench.BYTES_PER_CLIP == 4 * 66150 + 224 == 264824
    assert bench.FRAMES_PER_CLIP == 130
    assert sum(bench.FLOPS_PER_FRAME.values()) == 100883
    assert bench.FLOPS_PER_CLIP == 130 * 100883
    assert bench.FLOPS_PER_CLIP_FP32 == 130 * (100883 - 24600)            # chroma runs on the tensor cores
    assert 11.0e6 <= bench.FLOPS_PER_CLIP <= 13.6e6                         # SURVEY 8d's range
