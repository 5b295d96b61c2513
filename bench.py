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
def _oracle_worker(args):
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    seed, count = args
    import synth
    from oracle import librosa_port as lp
    w = synth.make_batch(count, N_SAMPLES, seed=seed)
    t0 = time.perf_counter()
    feats = lp.features_batch(w)
    return time.perf_counter() - t0, float(feats.sum())


def cpu_reference_rate(total_clips, cores):
    """Oracle port (librosa-equivalent restatement, 4 STFTs per clip like the reference) on `cores` processes.
    Each worker synthesises its own clips first (untimed) and times only the extraction; the rate is
    clips / max(worker extraction time)."""
    import multiprocessing as mp
    per = max(1, total_clips // cores)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        res = pool.map(_oracle_worker, [(1000 + i, per) for i in range(cores)])
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
    for _ in range(min(args.warmup, 1)):
        cpu_reference_rate(cores, cores)
    rates, tot, t0 = [], 0, time.perf_counter()
    for _ in range(args.steps):
        r, nclips, _ = cpu_reference_rate(per_step, cores)
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
        "cpu_baseline": {"value": value, "unit": "clips/s", "cores": cores, "kind": "port",
                         "sample": f"{tot} clips in {len(rates)} steps; oracle/librosa_port.py (librosa 0.10.0 "
                                   f"restatement, 4 STFTs per clip; real librosa is not installable here); {cpu_model()}"},
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


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "of measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "of fallback (6.65 TB/s, B200_PROFILING.md)"


def ncu_traffic_per_clip():
    """dram bytes per clip from the committed ncu capture of the same kernel (profiles/ncu_summary.json)."""
    try:
        return float(json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json")))["dram_bytes_per_clip"])
    except Exception:
        return None


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
    ap.add_argument("--e2e-clips", type=int, default=4096, help="clips per GPU per end-to-end (host buffer) step")
    ap.add_argument("--no-allgather", action="store_true")
    ap.add_argument("--allgather", choices=["overlap", "serial"], default="serial",
                    help="serial (default): the all-gather is waited for inside its step; overlap: step i's all-gather runs under "
                         "step i+1's extraction (its polling NCCL CTAs can delay the persistent extraction kernel)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
        rate, nclips, secs = cpu_reference_rate(total, cores)
        rate1, n1, secs1 = cpu_reference_rate(8, 1)
        cpu_baseline = {"value": rate, "unit": "clips/s", "cores": cores, "kind": "port",
                        "sample": f"{nclips} synthetic 3 s clips (tests/synth.py mix) over {cores} processes in {secs:.1f} s; "
                                  f"single core: {rate1:.1f} clips/s ({n1} clips); oracle/librosa_port.py = librosa 0.10.0 "
                                  f"restatement with the reference's 4 STFTs per clip (librosa itself is not installable "
                                  f"here); {cpu_model()}"}

    import numpy as np
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
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
    launches = ex.launches - launches0
    clocks = sampler.stop() if sampler else None
    if world > 1:
        tm = torch.tensor([total_ms, kern_ms], device=device, dtype=torch.float64)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        total_ms, kern_ms = tm.tolist()
    ms_per_step = total_ms / args.steps
    value = n_total / (ms_per_step * 1e-3)

    # ---------------- end-to-end: host (pinned) buffers through the C ABI's host entry point
    Be = min(args.e2e_clips, B)
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
    e2e_value = Be * world * e2e_steps / e2e_s
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
    e2e_pcm16 = {"value": Be * world * e2e_steps / pcm_s, "unit": "clips/s", "h2d_bytes_per_step": Be * N_SAMPLES * 2,
                 "d2h_bytes_per_step": Be * 56 * 4, "clips_per_gpu_per_step": Be, "steps": e2e_steps,
                 "path": "sfx_extract_host_pcm16: pinned int16 PCM rows -> H2D || device x/32768 + kernel || D2H",
                 "finite": bool(torch.isfinite(h_out16).all())}

    # file-shaped input: 3 s of 48 kHz mono 16-bit PCM per clip (what a RAVDESS WAV file holds), resampled to 22.05 kHz on
    # the device by the load_audio front-end (scope row f3), then extracted
    n48 = 48000 * 3
    Bf = min(Be, 2048)
    h_48 = torch.empty((Bf, n48), dtype=torch.int16).pin_memory()
    h_48.copy_((torch.randn((Bf, n48), generator=torch.Generator().manual_seed(5)) * 3000.0).round().clamp(-32768, 32767).to(torch.int16))
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

    # ---------------- parity spot check against the oracle on identical waveforms (rank 0)
    parity = None
    if rank == 0:
        import synth
        from oracle import librosa_port as lp
        idx = list(range(0, 16))
        w = pool[idx].cpu().numpy()
        ok, report = synth.compare(out[idx].cpu().numpy(), lp.features_batch(w))
        parity = {"ok": ok, "clips": len(idx), "e2e_bitwise_equal_device_path": e2e_match,
                  "tolerance": "|err| <= 1e-3*|ref| + atol(group) (tests/synth.py)",
                  "report": report.replace("\n", " | ")}

    if rank == 0:
        peak, peak_src = measured_peak()
        per_gpu = B / (kern_ms * 1e-3)
        achieved = per_gpu * BYTES_PER_CLIP / 1e9
        traffic = ncu_traffic_per_clip()
        line = {
            "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "clips_per_gpu_per_step": B, "n_samples": N_SAMPLES, "frames_per_clip": 1 + N_SAMPLES // 512,
                       "signal_mix": list(KINDS), "allgather_feature_cache": do_gather,
                       "allgather_overlap": (("step i's all-gather runs under step i+1's extraction" if args.allgather == "overlap"
                                             else "serial: waited for inside its step") if do_gather else None),
                       "l2": f"inputs larger than L2 ({B * N_SAMPLES * 4 / 1e9:.1f} GB per GPU per step)"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": Be * N_SAMPLES * 4,
                    "d2h_bytes_per_step": Be * 56 * 4, "clips_per_gpu_per_step": Be, "steps": e2e_steps,
                    "path": "sfx_extract_host: pinned host rows -> chunked H2D || kernel || D2H on 2 streams"},
            "e2e_pcm16": e2e_pcm16,
            "e2e_pcm16_48k": e2e_48k,
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (traffic * B) if traffic else None,
                         "note": f"{peak_src}; algorithmic bytes/launch = {B} clips x {BYTES_PER_CLIP} B; kernel avg "
                                 f"{kern_ms:.3f} ms (CUDA events); path is FP32-issue/SMEM bound (DESIGN.md), not HBM bound"},
            "roofline_fp32": fp32_roofline(ex, per_gpu),
            "cpu_baseline": cpu_baseline,
            "small_batch": small,
            "parity": parity,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
