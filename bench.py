#!/usr/bin/env python
"""bench.py -- throughput of the batched speech feature extractor (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo (sm_100a kernels through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port; librosa absent)

One "step" = one pass of the hot path over one resident batch of synthetic 3 s clips per GPU.  Workload =
BASELINE.json configs[3] ("1M synthetic 3 s clips sharded across 1/2/4/8 B200"): 1M x 264.6 KB does not fit one
GPU, so every GPU extracts a resident pool of --clips distinct clips (default 65 536 = 17.3 GB >> 126 MB L2) per
step; weak scaling (per-GPU work fixed); for N > 1 the step ends with the NCCL all-gather of the [N*B, 56] feature
cache.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, "multimodal-emotion-classification_b200"), ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

N_SAMPLES = 66150
BYTES_PER_CLIP = 4 * N_SAMPLES + 56 * 4          # algorithmic bytes: waveform in + 56 floats out
METRIC = "clips/sec (3 s @22.05 kHz -> 56-dim)"
# algorithmic flops per STFT frame (SURVEY 8d's itemisation, DESIGN.md "Roofline"); FMA = 2 flops
FLOPS_PER_FRAME = {
    "hann window": 2048,
    "1024-point complex FFT (5 N log2 N)": 51200,
    "real-FFT unpack (512 conjugate pairs x 14)": 7168,
    "|X|^2 and |X| (1025 bins x 4)": 4100,
    "mel projection (2018 non-zeros x 2)": 4036,
    "10 log10 (128 bands x 2)": 256,
    "piptrack (358 bins x 8)": 2864,
    "centroid and roll-off (1025 bins x 3)": 3075,
    "rms and zero crossings (512 samples x 3)": 1536,
    "chroma projection (12 x 1025 FMA)": 24600,       # runs on the tensor cores (FP16 hi/lo MMA), not on the FP32 pipes
}
FRAMES_PER_CLIP = 1 + N_SAMPLES // 512
FLOPS_PER_CLIP = FRAMES_PER_CLIP * sum(FLOPS_PER_FRAME.values())
FLOPS_PER_CLIP_FP32 = FLOPS_PER_CLIP - FRAMES_PER_CLIP * FLOPS_PER_FRAME["chroma projection (12 x 1025 FMA)"]
KINDS = ("noise", "harmonic", "noise_tail", "harmonic_tail")
WORKLOAD = "configs[3]: 1M x 3 s clips @22.05 kHz, clip-sharded; resident pool per GPU per step"


def source_sha16():
    """Hash of the kernel sources: profiles/ncu_summary.json carries the hash of the build it was captured from, so that the
    DRAM traffic reported as roofline.traffic can never silently belong to an older kernel."""
    import hashlib
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "multimodal-emotion-classification_b200", "csrc")
    for name in sorted(os.listdir(csrc)):
        if name.endswith((".cu", ".cuh", ".h")):
            h.update(name.encode())
            h.update(open(os.path.join(csrc, name), "rb").read())
    return h.hexdigest()[:16]


def librosa_available():
    try:
        import librosa  # noqa: F401
        return True
    except Exception:
        return False


def fp32_roofline(ex, clips_per_s_per_gpu):
    """Second, compute-side roofline: the algorithmic FP32-pipe flop rate of the extraction kernel against the
    FP32 FMA peak measured on this GPU by the library's register-only FFMA kernel (sfx_measure_fp32_peak)."""
    import ctypes
    tf = ctypes.c_double(0.0)
    from sfx_b200 import _lib
    rc = _lib.load_bench().sfx_measure_fp32_peak(ex.index, ctypes.byref(tf))
    if rc != 0 or tf.value <= 0.0:
        return None
    achieved = clips_per_s_per_gpu * FLOPS_PER_CLIP_FP32 / 1e12
    return {"bound": "fp32", "achieved": achieved, "peak": tf.value, "unit": "TFLOP/s", "frac": achieved / tf.value,
            "flops_per_clip_fp32": FLOPS_PER_CLIP_FP32, "flops_per_clip_total": FLOPS_PER_CLIP,
            "note": "peak = FFMA micro-benchmark measured in this run (CUDA events, best of 3); achieved = algorithmic flops "
                    "on the FP32 pipes (everything but the chroma contraction, which runs as FP16 hi/lo MMA) x clips/s of "
                    "the timed kernel; supplementary to the HBM roofline BASELINE.json names"}


# --------------------------------------------------------------------------------------------- CPU reference arm
def _reference_rows(w):
    """The reference's own three extract_* calls per clip (preprocessing/audio_preprocessing.py:22-37) on real librosa."""
    import librosa
    import numpy as np
    out = np.empty((len(w), 56), dtype=np.float32)
    for i, y in enumerate(w):
        mfcc = np.mean(librosa.feature.mfcc(y=y, sr=22050, n_mfcc=40).T, axis=0)
        chroma = np.mean(librosa.feature.chroma_stft(y=y, sr=22050).T, axis=0)
        spec = [float(np.mean(librosa.feature.zero_crossing_rate(y))), float(np.mean(librosa.feature.spectral_centroid(y=y, sr=22050))),
                float(np.mean(librosa.feature.spectral_rolloff(y=y, sr=22050))), float(np.mean(librosa.feature.rms(y=y)))]
        out[i] = np.concatenate([mfcc, chroma, np.array(spec, dtype=np.float32)])
    return out


def _oracle_worker(args):
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    seed, count, use_librosa = args
    import synth
    w = synth.make_batch(count, N_SAMPLES, seed=seed)
    if use_librosa:
        _reference_rows(w[:1])                                   # numba JIT warm-up outside the timed part
        t0 = time.perf_counter()
        feats = _reference_rows(w)
    else:
        from oracle import librosa_port as lp
        t0 = time.perf_counter()
        feats = lp.features_batch(w)
    return time.perf_counter() - t0, float(feats.sum())


def cpu_reference_rate(total_clips, cores, use_librosa=False):
    """The reference's CPU path on `cores` processes: real librosa through the reference's calls when it is importable, else
    the oracle port (librosa-equivalent restatement, 4 STFTs per clip like the reference).  Each worker synthesises its own
    clips first (untimed) and times only the extraction; the rate is clips / max(worker extraction time)."""
    import multiprocessing as mp
    per = max(1, total_clips // cores)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        res = pool.map(_oracle_worker, [(1000 + i, per, use_librosa) for i in range(cores)])
    slowest = max(r[0] for r in res)
    return per * cores / slowest, per * cores, slowest


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = host_cores()
    per_step = max(cores * 16, 16)
    real = librosa_available()
    for _ in range(min(args.warmup, 1)):
        cpu_reference_rate(cores, cores, real)
    rates, tot, t0 = [], 0, time.perf_counter()
    for _ in range(args.steps):
        r, nclips, _ = cpu_reference_rate(per_step, cores, real)
        rates.append(r)
        tot += nclips
        if time.perf_counter() - t0 > 150:
            break
    value = sum(rates) / len(rates)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": args.gpus,
        "steps": len(rates), "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * per_step / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64/f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_samples": N_SAMPLES, "frames_per_clip": 1 + N_SAMPLES // 512,
                   "signal_mix": list(KINDS), "clips_per_step": per_step,
                   "sample": f"bounded sample of that workload: {per_step} clips per step, same synthetic distributions "
                             "(tests/synth.py), all host cores"},
        "cpu_baseline": {"value": value, "unit": "clips/s", "cores": cores, "kind": "reference" if real else "port",
                         "sample": (f"{tot} clips in {len(rates)} steps; " +
                                    ("librosa itself through the reference's extract_* calls; " if real else
                                     "oracle/librosa_port.py (librosa 0.10.0 restatement, 4 STFTs per clip; real librosa is "
                                     "not importable here); ") + cpu_model())},
        "e2e": {"value": value, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# --------------------------------------------------------------------------------------------- synthetic pool on device
def synth_pool(B, n, seed, device, chunk=2048):
    """Device-side synthetic clips with the distributions of tests/synth.py (noise, harmonic stacks, zeroed tails)."""
    import math
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    out = torch.empty((B, n), dtype=torch.float32, device=device)
    t = torch.arange(n, device=device, dtype=torch.float32) / 22050.0
    for c0 in range(0, B, chunk):
        nb = min(chunk, B - c0)
        blk = out[c0:c0 + nb]
        idx = torch.arange(c0, c0 + nb, device=device)
        kind = idx % 4
        u = torch.rand((nb, 8), device=device, generator=g)
        blk.normal_(0.0, 0.1, generator=g)                                          # noise rows (and noise floor)
        harm = (kind % 2 == 1).nonzero().squeeze(1)
        if harm.numel():
            uh = u[harm]
            f0 = 90.0 + 210.0 * uh[:, 0:1]
            glide = (uh[:, 1:2] - 0.5) * 0.3 * f0
            vib = uh[:, 2:3] * 0.03 * f0
            fv = 4.0 + 3.0 * uh[:, 3:4]
            inst = f0 + glide * (t / t[-1]) + vib * torch.sin(2 * math.pi * fv * t)
            phase = 2 * math.pi * torch.cumsum(inst.double(), dim=1).float() / 22050.0
            y = torch.zeros_like(phase)
            for h in range(1, 23):
                y += torch.sin(h * phase + 6.2831853 * uh[:, 4:5] * h) / h
            env = 0.55 + 0.45 * torch.sin(2 * math.pi * (1.5 + 2.5 * uh[:, 5:6]) * t + 6.2831853 * uh[:, 6:7])
            y = y * env
            y = 0.5 * y / y.abs().amax(dim=1, keepdim=True).clamp_min(1e-9)
            blk[harm] = y + blk[harm] * (10 ** (-50 / 20) * 0.5 / 0.1)
        tail = (kind >= 2).nonzero().squeeze(1)
        if tail.numel():
            cut = ((0.55 + 0.25 * u[tail, 7]) * n).long()
            mask = torch.arange(n, device=device)[None, :] >= cut[:, None]
            blk[tail] = blk[tail].masked_fill(mask, 0.0)
        blk.clamp_(-1.0, 1.0)
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            txt = self.proc.communicate(timeout=5)[0]
        except subprocess.TimeoutExpired:
            self.proc.kill()
            txt = ""
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in txt.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "samples": len(sm),
                "reasons": sorted(reasons)}


def bind_to_gpu_numa(index):
    """Pin this rank's host threads to the CPUs local to its GPU (/sys/bus/pci/devices/<bdf>/local_cpulist) BEFORE any pinned
    host buffer is allocated, so that first-touch places the staging memory on the GPU's NUMA node.  Returns what was found."""
    info = {"numa_node": None, "cpus_bound": None}
    try:
        import torch
        pr = torch.cuda.get_device_properties(index)
        bdf = f"{getattr(pr, 'pci_domain_id', 0):04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bdf}"
        info["pci"] = bdf
        node = int(open(base + "/numa_node").read().strip())
        info["numa_node"] = node
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            if not part:
                continue
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        use = sorted(cpus & allowed)
        if node >= 0 and use and len(use) < len(allowed):
            os.sched_setaffinity(0, use)
            info["cpus_bound"] = f"{use[0]}-{use[-1]} ({len(use)})"
    except Exception as e:      # noqa: BLE001
        info["note"] = f"not bound: {e}"
    return info


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "of measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "of fallback (6.65 TB/s, B200_PROFILING.md)"


def ncu_traffic_per_clip():
    """(dram bytes per clip, note) from the committed ncu capture (profiles/ncu_summary.json) -- only if that capture was made
    from the kernel sources this run is built from (its src_sha16 equals source_sha16()); a stale capture reports None."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json")))
        if d.get("src_sha16") != source_sha16():
            return None, f"profiles/ncu_summary.json is from sources {d.get('src_sha16')}, this build is {source_sha16()}: not reported"
        return float(d["dram_bytes_per_clip"]), f"dram__bytes_read+write per clip from {d.get('source')}"
    except Exception as e:      # noqa: BLE001
        return None, f"no ncu summary ({e})"


# --------------------------------------------------------------------------------------------- main arm
def emit(line: dict):
    """The ONE JSON line goes to the process's original stdout; everything else (NCCL banner, library chatter)
    was redirected to stderr at the file-descriptor level."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = os.dup(1)


def main():
    os.dup2(2, 1)                  # C-level writers (e.g. "NCCL version ...") must not precede the JSON line on stdout
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--clips", type=int, default=65536, help="resident clips per GPU per step")
    ap.add_argument("--e2e-clips", type=int, default=16384,
                    help="clips per GPU per end-to-end (host buffer) step (halved until its pinned rows fit a share of the host's free memory)")
    ap.add_argument("--no-allgather", action="store_true")
    ap.add_argument("--allgather", choices=["overlap", "serial"], default="serial",
                    help="serial (default): the all-gather is waited for inside its step; overlap: step i's all-gather runs under "
                         "step i+1's extraction (its polling NCCL CTAs can delay the persistent extraction kernel)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config5", action="store_true", help="skip the 4 096-clip 0.5-60 s batch (needs 22 GB of device memory)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    # CPU baseline first (fork pool before CUDA is initialised): rank 0, N = 1 only
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        total = min(4096, max(64, 48 * cores))          # ~15-25 s of CPU work spread over the cores
        real = librosa_available()
        rate, nclips, secs = cpu_reference_rate(total, cores, real)
        rate1, n1, secs1 = cpu_reference_rate(8, 1, real)
        cpu_baseline = {"value": rate, "unit": "clips/s", "cores": cores, "kind": "reference" if real else "port",
                        "sample": f"{nclips} synthetic 3 s clips (tests/synth.py mix) over {cores} processes in {secs:.1f} s; "
                                  f"single core: {rate1:.1f} clips/s ({n1} clips); " +
                                  ("librosa itself through the reference's extract_* calls; " if real else
                                   "oracle/librosa_port.py = librosa 0.10.0 restatement with the reference's 4 STFTs per clip "
                                   "(librosa itself is not importable here); ") + cpu_model()}

    import numpy as np
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    numa = bind_to_gpu_numa(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    from sfx_b200 import get_extractor
    from sfx_b200.shard import gather_feature_cache
    ex = get_extractor(device)
    B = args.clips
    pool = synth_pool(B, N_SAMPLES, seed=1234 + rank, device=device)
    out = torch.empty((B, 56), dtype=torch.float32, device=device)
    n_total = B * world
    do_gather = world > 1 and not args.no_allgather
    # the all-gather of step i runs on the communication stream underneath the extraction of step i+1: two feature
    # buffers and two cache buffers, each reused only after its collective has been waited for
    outs = [out, torch.empty_like(out)] if do_gather else [out]
    caches = [torch.empty((n_total, 56), dtype=torch.float32, device=device) for _ in range(2)] if do_gather else []
    works = [None, None]

    def step(i):
        b = i & 1 if do_gather else 0
        if do_gather and works[b] is not None:
            works[b].wait()
        ex.extract(pool, out=outs[b])
        if do_gather:
            _, works[b] = gather_feature_cache(outs[b], n_total, out=caches[b], async_op=True)
            if args.allgather == "serial":
                works[b].wait()

    def drain():
        for b in range(2):
            if works[b] is not None:
                works[b].wait()
                works[b] = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    drain()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = ex.launches
    k0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    k1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        b = i & 1 if do_gather else 0
        if do_gather and works[b] is not None:
            works[b].wait()
        k0[i].record()
        ex.extract(pool, out=outs[b])
        k1[i].record()
        if do_gather:
            _, works[b] = gather_feature_cache(outs[b], n_total, out=caches[b], async_op=True)
            if args.allgather == "serial":
                works[b].wait()
    drain()                                   # the timed region ends when every step's cache is complete
    e1.record()
    barrier()
    total_ms = e0.elapsed_time(e1)
    kern_ms = sum(a.elapsed_time(b) for a, b in zip(k0, k1)) / args.steps
    # all-gather correctness on NCCL: on every rank, block r of the gathered feature cache must be
    # rank r's rows bit for bit.  Each rank publishes a digest of its rows (exact integer sums of the float bit patterns per
    # column, and of bits * (row index + 1)); every rank recomputes the W digests from its copy of the cache.
    allgather_verified = None
    if do_gather:
        def digest(rows):
            bits = rows.contiguous().view(torch.int32).to(torch.int64)
            k = torch.arange(1, rows.shape[0] + 1, device=rows.device, dtype=torch.int64)[:, None]
            return torch.cat([bits.sum(dim=0), (bits * k).sum(dim=0)])
        last = (args.steps - 1) & 1
        mine = digest(outs[last])
        every = torch.empty((world, mine.numel()), dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(every, mine)
        bad = sum(int(not torch.equal(digest(caches[last][r * B:(r + 1) * B]), every[r])) for r in range(world))
        bad += int(not torch.equal(caches[last][rank * B:(rank + 1) * B], outs[last]))
        badt = torch.tensor([bad], device=device, dtype=torch.int64)
        dist.all_reduce(badt, op=dist.ReduceOp.SUM)
        allgather_verified = int(badt.item()) == 0
    launches = ex.launches - launches0
    clocks = sampler.stop() if sampler else None
    if world > 1:
        tm = torch.tensor([total_ms, kern_ms], device=device, dtype=torch.float64)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        total_ms, kern_ms = tm.tolist()
    ms_per_step = total_ms / args.steps
    value = n_total / (ms_per_step * 1e-3)

    # ---------------- end-to-end: host (pinned) buffers through the C ABI's host entry point
    # (a call's last chunk cannot overlap its kernel and D2H with a copy: ~0.6 ms per call, 6 % of a 4 096-clip step from
    #  PCM rows, 1.5 % of a 16 384-clip one.  The step's pinned rows -- float32 + int16 -- take 6 bytes per sample per rank.)
    Be = min(args.e2e_clips, B)
    try:
        import psutil
        host_share = psutil.virtual_memory().available // (4 * max(1, int(os.environ.get("LOCAL_WORLD_SIZE", world))))
        while Be > 1024 and Be * N_SAMPLES * 6 > host_share:
            Be //= 2
    except ImportError:
        pass
    if world > 1:                      # one size on every rank (the aggregate is Be * world clips per step)
        tb_ = torch.tensor([Be], device=device, dtype=torch.int64)
        dist.all_reduce(tb_, op=dist.ReduceOp.MIN)
        Be = int(tb_.item())
    h_in = torch.empty((Be, N_SAMPLES), dtype=torch.float32).pin_memory()
    h_in.copy_(pool[:Be])
    h_out = torch.empty((Be, 56), dtype=torch.float32).pin_memory()
    e2e_steps = max(3, min(args.steps, 8))
    for _ in range(2):
        ex.extract_host(h_in.numpy(), out=h_out.numpy())
    barrier()
    launches1 = ex.launches
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ex.extract_host(h_in.numpy(), out=h_out.numpy())
    e2e_s = time.perf_counter() - t0
    launches += ex.launches - launches1
    if world > 1:
        tm = torch.tensor([e2e_s], device=device, dtype=torch.float64)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        e2e_s = tm.item()
    e2e_f32_value = Be * world * e2e_steps / e2e_s
    e2e_match = bool(torch.equal(h_out, out[:Be].cpu()))

    # the same clips as 16-bit PCM (what a WAV file at 22.05 kHz holds): half the bytes over PCIe, dequantised on the
    # device exactly as soundfile does for librosa.load; reported beside `e2e`, not instead of it
    h_pcm = torch.empty((Be, N_SAMPLES), dtype=torch.int16).pin_memory()
    h_pcm.copy_((pool[:Be] * 32768.0).round().clamp(-32768, 32767).to(torch.int16))
    h_out16 = torch.empty((Be, 56), dtype=torch.float32).pin_memory()
    for _ in range(2):
        ex.extract_host(h_pcm.numpy(), out=h_out16.numpy())
    barrier()
    launches1 = ex.launches
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ex.extract_host(h_pcm.numpy(), out=h_out16.numpy())
    pcm_s = time.perf_counter() - t0
    launches += ex.launches - launches1
    if world > 1:
        tm = torch.tensor([pcm_s], device=device, dtype=torch.float64)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        pcm_s = tm.item()
    # the 16-bit rows are what a 22.05 kHz WAV file holds (the reference's input is files, reference :13): this is `e2e`;
    # the float32 rows (the buffers the reference's extract_* functions are handed) are reported as `e2e_f32`
    deq = (h_pcm[:64].to(torch.float32) / 32768.0).to(device)
    pcm_bitwise = bool(torch.equal(h_out16[:64], ex.extract(deq).cpu()))
    e2e = {"value": Be * world * e2e_steps / pcm_s, "unit": "clips/s", "h2d_bytes_per_step": Be * N_SAMPLES * 2,
           "d2h_bytes_per_step": Be * 56 * 4, "clips_per_gpu_per_step": Be, "steps": e2e_steps,
           "h2d_gb_per_s_per_gpu": Be * e2e_steps * N_SAMPLES * 2 / pcm_s / 1e9, "host_binding_rank0": numa,
           "input": "16-bit PCM rows as a 22.05 kHz mono WAV file holds them (pinned host memory)",
           "path": "sfx_extract_host_pcm16: pinned int16 PCM rows -> chunked H2D || device x/32768 (soundfile's conversion) + "
                   "kernel || D2H of the 56-float rows",
           "finite": bool(torch.isfinite(h_out16).all()), "rows_bitwise_equal_float_path_on_dequantised_samples": pcm_bitwise}
    e2e_f32 = {"value": e2e_f32_value, "unit": "clips/s", "h2d_bytes_per_step": Be * N_SAMPLES * 4,
               "d2h_bytes_per_step": Be * 56 * 4, "clips_per_gpu_per_step": Be, "steps": e2e_steps,
               "h2d_gb_per_s_per_gpu": Be * e2e_steps * N_SAMPLES * 4 / e2e_s / 1e9,
               "path": "sfx_extract_host: pinned float32 rows -> chunked H2D || kernel || D2H on 3 streams",
               "rows_bitwise_equal_device_path": e2e_match}

    # file-shaped input: 3 s of 48 kHz mono 16-bit PCM per clip (what a RAVDESS WAV file holds), resampled to 22.05 kHz on
    # the device by the load_audio front-end (scope row f3), then extracted
    n48 = 48000 * 3
    Bf = min(Be, 8192)
    h_48 = torch.empty((Bf, n48), dtype=torch.int16).pin_memory()
    base48 = (torch.randn((min(Bf, 2048), n48), generator=torch.Generator().manual_seed(5)) * 3000.0).round().clamp(-32768, 32767).to(torch.int16)
    for r0 in range(0, Bf, base48.shape[0]):          # 2 048 distinct clips, repeated
        h_48[r0:r0 + base48.shape[0]].copy_(base48[:Bf - r0])
    del base48
    h_out48 = torch.empty((Bf, 56), dtype=torch.float32).pin_memory()
    for _ in range(2):
        ex.preprocess_pcm16(h_48.numpy(), None, 48000, out=h_out48.numpy())
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ex.preprocess_pcm16(h_48.numpy(), None, 48000, out=h_out48.numpy())
    f_s = time.perf_counter() - t0
    if world > 1:
        tm = torch.tensor([f_s], device=device, dtype=torch.float64)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        f_s = tm.item()
    e2e_48k = {"value": Bf * world * e2e_steps / f_s, "unit": "clips/s", "h2d_bytes_per_step": Bf * n48 * 2,
               "d2h_bytes_per_step": Bf * 56 * 4, "clips_per_gpu_per_step": Bf, "steps": e2e_steps,
               "path": "sfx_preprocess_host_pcm16: pinned 48 kHz int16 PCM rows -> H2D || x/32768 + polyphase resample to "
                       "22.05 kHz (float64, scipy.resample_poly-identical) + extractor || D2H",
               "finite": bool(torch.isfinite(h_out48).all())}

    # ---------------- small-batch behaviour (rank 0): single-clip latency through the host path (what one request of the
    # reference's Flask app costs), and device-resident throughput at the batch sizes of configs[0] / configs[1]
    small = None
    if rank == 0:
        small = {}
        one_in = h_in[:1].numpy()
        one_out = h_out[:1].numpy()
        for _ in range(5):
            ex.extract_host(one_in, out=one_out)
        t0 = time.perf_counter()
        for _ in range(50):
            ex.extract_host(one_in, out=one_out)
        small["single_clip_host_latency_ms"] = (time.perf_counter() - t0) / 50 * 1e3
        for nb, tag in ((64, "config1_64_clips"), (1440, "config2_1440_clips")):
            sub = pool[:nb]
            sub_out = out[:nb]
            for _ in range(3):
                ex.extract(sub, out=sub_out)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            for _ in range(20):
                ex.extract(sub, out=sub_out)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 20
            small[tag] = {"ms_per_batch": ms, "clips_per_s": nb / (ms * 1e-3)}
        ex.extract(pool, out=out)          # restore the full-batch output for the parity check below
        torch.cuda.synchronize()

    # ---------------- BASELINE configs[2] (rank 0): 2 800 TESS-shaped clips (~2 s zero-padded to 3 s) -> features -> scaler ->
    # speech DNN, everything device-resident (scope row f1); synthetic weights of the reference architecture
    config3 = None
    if rank == 0:
        from sfx_b200.dnn import SpeechDNN
        rs = np.random.default_rng(3)
        widths = [56, 512, 512, 256, 128, 64, 7]
        wts = {"widths": widths}
        for i in range(6):
            wts[f"kernel{i}"] = (rs.standard_normal((widths[i], widths[i + 1])) * np.sqrt(2.0 / widths[i])).astype(np.float32)
            wts[f"bias{i}"] = (0.01 * rs.standard_normal(widths[i + 1])).astype(np.float32)
            if i < 5:
                wts[f"gamma{i}"] = (1.0 + 0.1 * rs.standard_normal(widths[i + 1])).astype(np.float32)
                wts[f"beta{i}"] = (0.1 * rs.standard_normal(widths[i + 1])).astype(np.float32)
                wts[f"mean{i}"] = (0.1 * rs.standard_normal(widths[i + 1])).astype(np.float32)
                wts[f"var{i}"] = (1.0 + 0.1 * rs.random(widths[i + 1])).astype(np.float32)
        nb3 = min(2800, B)
        w3 = pool[:nb3].clone()
        cut = torch.from_numpy((44100 * rs.uniform(0.85, 1.15, nb3)).astype(np.int64)).to(device)
        w3.masked_fill_(torch.arange(N_SAMPLES, device=device)[None, :] >= cut[:, None], 0.0)
        f3 = ex.extract(w3)
        wts["scaler_mean"] = f3.double().mean(dim=0).cpu().numpy()
        wts["scaler_scale"] = f3.double().std(dim=0).clamp_min(1e-6).cpu().numpy()
        dnn = SpeechDNN(wts, device)
        for _ in range(3):
            probs3, _ = dnn.forward(ex.extract(w3))
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        torch.cuda.synchronize()
        a.record()
        for _ in range(10):
            f3 = ex.extract(w3)
        b.record()
        for _ in range(10):
            probs3, _ = dnn.forward(f3)
        c.record()
        torch.cuda.synchronize()
        ms_x, ms_d = a.elapsed_time(b) / 10, b.elapsed_time(c) / 10
        config3 = {"clips": nb3, "extract_ms": ms_x, "scaler_dnn_forward_ms": ms_d, "clips_per_s": nb3 / ((ms_x + ms_d) * 1e-3),
                   "dnn_launches_per_forward": dnn.launches // 13, "probabilities_sum_to_one": bool(
                       torch.allclose(probs3.sum(dim=1), torch.ones(nb3, device=device), atol=1e-5)),
                   "note": "features never leave the device; weights are synthetic (the reference's .h5 needs TF/h5py)"}
        del w3, f3

    # ---------------- BASELINE configs[4] (rank 0): 4 096 clips, lengths log-uniform in [0.5 s, 60 s], padded rows + lengths
    config5 = None
    if rank == 0 and not args.no_config5:
        rs = np.random.default_rng(5)
        B5, n_min, n_max = 4096, 11025, 1323000
        lens5 = np.exp(rs.uniform(np.log(n_min), np.log(n_max), size=B5)).astype(np.int64)
        lens5[0], lens5[-1] = n_min, n_max
        w5 = torch.empty((B5, n_max), dtype=torch.float32, device=device)
        g5 = torch.Generator(device=device).manual_seed(5)
        for c0 in range(0, B5, 512):
            w5[c0:c0 + 512].normal_(0.0, 0.1, generator=g5)
        ld5 = torch.from_numpy(lens5.astype(np.int32)).to(device)
        o5 = torch.empty((B5, 56), dtype=torch.float32, device=device)
        for _ in range(2):
            ex.extract(w5, ld5, out=o5)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(3):
            ex.extract(w5, ld5, out=o5)
        b.record()
        torch.cuda.synchronize()
        ms5 = a.elapsed_time(b) / 3
        frames5 = int((1 + lens5 // 512).sum())
        # the same kind of signal (white noise) at the fixed 3 s length, for the ratio
        wn = pool[0:B:4][:8192].contiguous()
        on = torch.empty((wn.shape[0], 56), dtype=torch.float32, device=device)
        for _ in range(2):
            ex.extract(wn, out=on)
        torch.cuda.synchronize()
        a.record()
        for _ in range(3):
            ex.extract(wn, out=on)
        b.record()
        torch.cuda.synchronize()
        fixed_frames_per_s = wn.shape[0] * FRAMES_PER_CLIP / (a.elapsed_time(b) / 3 * 1e-3)
        import synth
        import oracle_pool
        pick = np.concatenate([[0], np.argsort(lens5)[1:12], rs.choice(B5, size=4, replace=False)])
        ok5, rep5 = synth.compare(o5[torch.from_numpy(pick).to(device)].cpu().numpy(),
                                  oracle_pool.oracle_rows(w5[torch.from_numpy(pick).to(device)].cpu().numpy(), lens5[pick], ctx="spawn"))
        config5 = {"clips": B5, "seconds_of_audio": float(lens5.sum() / 22050.0), "frames": frames5, "ms_per_batch": ms5,
                   "clips_per_s": B5 / (ms5 * 1e-3), "frames_per_s": frames5 / (ms5 * 1e-3),
                   "frames_per_s_fixed_3s_noise": fixed_frames_per_s,
                   "ragged_over_fixed": frames5 / (ms5 * 1e-3) / fixed_frames_per_s, "finite": bool(torch.isfinite(o5).all()),
                   "parity_ok": ok5, "parity_clips": int(len(pick)), "signal": "white noise (the worst case for the peak list)"}
        del w5, o5

    # ---------------- parity against the oracle on identical waveforms (rank 0): 256 clips strided over the whole pool
    parity = None
    if rank == 0:
        import synth
        import oracle_pool
        idx = np.arange(0, B, max(1, B // 256))[:256]
        it = torch.from_numpy(idx).to(device)
        w = pool[it].cpu().numpy()
        got = out[it].cpu().numpy()
        ref = oracle_pool.oracle_rows(w, ctx="spawn")
        ok, report = synth.compare(got, ref)
        rel = np.abs(got.astype(np.float64) - ref) / np.maximum(np.abs(ref.astype(np.float64)), 1e-30)
        parity = {"ok": ok, "clips": int(len(idx)), "e2e_f32_bitwise_equal_device_path": e2e_match,
                  "tolerance": "|err| <= 1e-3*|ref| + atol(group) (tests/synth.py)",
                  "pure_relative_error": {"median": float(np.median(rel)), "p99": float(np.quantile(rel, 0.99)),
                                          "p999": float(np.quantile(rel, 0.999)), "max": float(rel.max()),
                                          "fraction_above_1e-3": float((rel > 1e-3).mean()),
                                          "note": "north_star's pure relative form, no absolute term: the tail is pooled MFCCs "
                                                  "whose reference value is near zero"},
                  "report": report.replace("\n", " | ")}

    if rank == 0:
        peak, peak_src = measured_peak()
        per_gpu = B / (kern_ms * 1e-3)
        achieved = per_gpu * BYTES_PER_CLIP / 1e9
        traffic, traffic_note = ncu_traffic_per_clip()
        line = {
            "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "clips_per_gpu_per_step": B, "n_samples": N_SAMPLES, "frames_per_clip": 1 + N_SAMPLES // 512,
                       "signal_mix": list(KINDS), "allgather_feature_cache": do_gather,
                       "allgather_overlap": (("step i's all-gather runs under step i+1's extraction" if args.allgather == "overlap"
                                             else "serial: waited for inside its step") if do_gather else None),
                       "allgather_verified": allgather_verified,
                       "l2": f"inputs larger than L2 ({B * N_SAMPLES * 4 / 1e9:.1f} GB per GPU per step)"},
            "clocks": clocks,
            "e2e": e2e,
            "e2e_f32": e2e_f32,
            "e2e_pcm16_48k": e2e_48k,
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (traffic * B) if traffic else None, "traffic_note": traffic_note,
                         "note": f"{peak_src}; algorithmic bytes/launch = {B} clips x {BYTES_PER_CLIP} B; kernel avg "
                                 f"{kern_ms:.3f} ms (CUDA events); path is FP32-issue/SMEM bound (DESIGN.md), not HBM bound"},
            "roofline_fp32": fp32_roofline(ex, per_gpu),
            "cpu_baseline": cpu_baseline,
            "small_batch": small,
            "config3_tess_dnn": config3,
            "config5_ragged": config5,
            "parity": parity,
            "pipeline": os.environ.get("SFX_PIPELINE", "auto"),
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
