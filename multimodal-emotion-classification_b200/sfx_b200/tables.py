"""Host-side constant tables for the sm_100a speech feature extractor.

Everything is generated in float64 with numpy and rounded once to the dtype the kernels read, so the
filterbanks are the same numbers librosa 0.10.0 builds for the reference's calls
(audio_preprocessing.py:23 ``feature.mfcc``, :28 ``feature.chroma_stft``): ``filters.mel`` (Slaney,
128 bands, norm='slaney'), ``filters.chroma`` for each of the 100 tunings ``estimate_tuning`` can return,
the ortho DCT-II matrix of ``scipy.fftpack.dct``, the periodic Hann window of ``scipy.signal.get_window``,
and the histogram edges of ``pitch_tuning``.  FFT twiddles are cos/sin evaluated in float64.

Layouts are chosen for the kernels (see DESIGN.md "HBM / SMEM layout"), not for readability:
  hann   float32[2048]        (w[2m], w[2m+1]) pairs, read as float2[1024]
  tw1    float32[32][32][2]   tw1[k1][lane] = exp(-2*pi*i*lane*k1/1024)      (inter-stage twiddle)
  tw2    float32[32][32][2]   tw2[k2][lane] = 0.5*(cos, sin)(2*pi*(lane+32*k2)/2048)  (real-FFT unpack)
  mel_ab float32[33][32][2]   (falling, rising) Slaney weights of bin 32*lane + j at [j][lane]; mel_mask uint32[32]
                              flush bits, mel_src int32[128][3] partial-sum slots (see mel_chunk_layout)
  melw   float32[mel_rows][32] transposed/padded sparse mel weights (dense-bank audit layout, host tests only)
  chroma16 float16[100][2][12][1056] one bank per tuning edge as hi + 2^-11 * lo (operands of the tensor-core chroma
                              projection), bin axis zero-padded to 1056; chroma_ny float32[100][12] Nyquist-bin weights;
  chroma_frag uint32[100][32][2][2][32][4] the same bank as ready-made mma.m16n8k16 A fragments (stream pipeline);
                              chroma_f32 is the dense float32 bank (host tests only)
  dct    float64[128][128]    rows k of the ortho DCT-II
  edges  float64[101]         np.linspace(-0.5, 0.5, 101)
"""
from __future__ import annotations

import functools
import numpy as np

N_FFT = 2048
HOP = 512
N_BINS = 1025
N_MELS = 128
N_CHROMA = 12
N_TUNINGS = 100
P_STRIDE = 1056          # padded bin row (33 * 32)


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    logstep = np.log(6.4) / 27.0
    lin = f / f_sp
    with np.errstate(divide="ignore"):
        log = 1000.0 / f_sp + np.log(np.maximum(f, 1e-300) / 1000.0) / logstep
    return np.where(f >= 1000.0, log, lin)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    logstep = np.log(6.4) / 27.0
    min_log_mel = 1000.0 / f_sp
    return np.where(m >= min_log_mel, 1000.0 * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_bank(sr: int) -> np.ndarray:
    """Dense [128, 1025] float32 Slaney mel bank (librosa.filters.mel defaults used by feature.mfcc)."""
    fftfreqs = np.fft.rfftfreq(N_FFT, 1.0 / sr)
    edges_hz = _mel_to_hz(np.linspace(_hz_to_mel(0.0), _hz_to_mel(sr / 2.0), N_MELS + 2))
    width = np.diff(edges_hz)
    ramps = edges_hz[:, None] - fftfreqs[None, :]
    tri = np.maximum(0.0, np.minimum(-ramps[:-2] / width[:-1, None], ramps[2:] / width[1:, None]))
    tri = tri.astype(np.float32)                      # librosa stores the triangle in float32 first
    enorm = 2.0 / (edges_hz[2:] - edges_hz[:-2])
    return (tri.astype(np.float64) * enorm[:, None]).astype(np.float32)


def chroma_bank(sr: int, tuning: float) -> np.ndarray:
    """Dense [12, 1025] float32 chroma bank (librosa.filters.chroma: ctroct=5, octwidth=2, norm=2, base_c)."""
    freqs = np.linspace(0, sr, N_FFT, endpoint=False)[1:]
    a440 = 440.0 * 2.0 ** (tuning / N_CHROMA)
    frq = N_CHROMA * np.log2(freqs / (float(a440) / 16))
    frq = np.concatenate(([frq[0] - 1.5 * N_CHROMA], frq))
    bw = np.concatenate((np.maximum(frq[1:] - frq[:-1], 1.0), [1]))
    D = np.subtract.outer(frq, np.arange(0, N_CHROMA, dtype="d")).T
    half = np.round(float(N_CHROMA) / 2)
    D = np.remainder(D + half + 10 * N_CHROMA, N_CHROMA) - half
    w = np.exp(-0.5 * (2 * D / np.tile(bw, (N_CHROMA, 1))) ** 2)
    length = np.sum(np.abs(w) ** 2, axis=0, keepdims=True) ** 0.5
    length[length < np.finfo(np.float64).tiny] = 1.0
    w = w / length
    w *= np.tile(np.exp(-0.5 * (((frq / N_CHROMA - 5.0) / 2) ** 2)), (N_CHROMA, 1))
    w = np.roll(w, -3 * (N_CHROMA // 12), axis=0)
    return np.ascontiguousarray(w[:, :N_BINS], dtype=np.float32)


def round_to_tf32(x: np.ndarray) -> np.ndarray:
    """Round float32 to the TF32 grid (10-bit mantissa), nearest / ties away: what cvt.rna.tf32.f32 does."""
    b = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    return ((b + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def split_fp16(x: np.ndarray):
    """x ~= hi + lo * 2^-11 with hi, lo in float16 (lo is pre-scaled by 2^11 so it stays in the normal range)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    hi = x.astype(np.float16)
    lo = ((x.astype(np.float64) - hi.astype(np.float64)) * 2048.0).astype(np.float16)
    return hi, lo


def chroma_fragments(c_hi: np.ndarray, c_lo: np.ndarray) -> np.ndarray:
    """A-operand fragments of mma.m16n8k16 for the chroma bank, one 16-byte quad per (tuning, 32-bin step, half step,
    hi/lo, lane): uint32 [T][32][2][2][32][4].  Lane (g = lane / 4, t4 = lane % 4) of half step m of step s owns bins
    b0 = 32 s + 8 t4 + 4 m .. b0 + 3 (the same K permutation the kernels apply to the |X|^2 operand) and rows g, g + 8:
    quad = (W[g][b0:b0+2], W[g+8][b0:b0+2], W[g][b0+2:b0+4], W[g+8][b0+2:b0+4]) as packed half2; rows 12..15 are zero.
    One coalesced 128-bit load per MMA instead of two strided loads and four register moves."""
    nt = c_hi.shape[0]
    a = np.zeros((nt, 2, 16, 1024), dtype=np.uint16)
    a[:, 0, :N_CHROMA] = np.ascontiguousarray(c_hi[:, :, :1024]).view(np.uint16)
    a[:, 1, :N_CHROMA] = np.ascontiguousarray(c_lo[:, :, :1024]).view(np.uint16)
    a = a.reshape(nt, 2, 2, 8, 32, 4, 2, 2, 2).astype(np.uint32)          # [tun, hl, rh, g, s, t4, m, jp, e]
    u = a[..., 0] | (a[..., 1] << np.uint32(16))                             # [tun, hl, rh, g, s, t4, m, jp]
    u = u.transpose(0, 4, 6, 1, 3, 5, 7, 2)                                  # [tun, s, m, hl, g, t4, jp, rh]
    return np.ascontiguousarray(u.reshape(nt, 32, 2, 2, 32, 4))


def chroma_umma_images(c_hi: np.ndarray, c_lo: np.ndarray) -> np.ndarray:
    """B operand of the tcgen05 chroma projection (csrc: SFX_CHROMA_UMMA), one shared-memory image per (tuning, 64-bin K block):
    32 rows (0..11 = hi of chroma class c, 16..27 = 2^11 * lo, the rest zero) x 64 FP16 bins, K-major with the 128-byte swizzle
    the UMMA shared-memory descriptor expects (8-row x 128-byte atoms, 16-byte chunk index XOR row-in-atom), so that one bulk
    copy of 4 096 contiguous bytes lands it ready to use.  uint8 [T][16][4096]."""
    nt = c_hi.shape[0]
    rows = np.zeros((nt, 32, 1024), dtype=np.uint16)
    rows[:, :N_CHROMA] = np.ascontiguousarray(c_hi[:, :, :1024]).view(np.uint16)
    rows[:, 16:16 + N_CHROMA] = np.ascontiguousarray(c_lo[:, :, :1024]).view(np.uint16)
    blk = rows.reshape(nt, 32, 16, 8, 8)                       # [tun, n, kb, chunk, half-in-chunk]
    n = np.arange(32)
    out = np.zeros((nt, 16, 4, 8, 8, 8), dtype=np.uint16)      # [tun, kb, atom, row-in-atom, chunk position, half]
    for c in range(8):
        pos = c ^ (n & 7)
        out[:, :, n >> 3, n & 7, pos, :] = blk[:, :, :, c, :].transpose(0, 2, 1, 3)
    return np.ascontiguousarray(out).view(np.uint8).reshape(nt, 16, 4096)


def tuning_edges() -> np.ndarray:
    return np.linspace(-0.5, 0.5, N_TUNINGS + 1)


def dct_matrix() -> np.ndarray:
    """[128,128] float64: row k of scipy.fftpack.dct(type=2, norm='ortho') over 128 mel bands."""
    m = np.arange(N_MELS, dtype=np.float64)
    k = np.arange(N_MELS, dtype=np.float64)[:, None]
    d = np.cos(np.pi * k * (2.0 * m + 1.0) / (2.0 * N_MELS)) * np.sqrt(2.0 / N_MELS)
    d[0] = 1.0 / np.sqrt(N_MELS)
    return np.ascontiguousarray(d)


def piptrack_bin_range(sr: int, fmin: float = 150.0, fmax: float = 4000.0):
    """First/last rFFT bin with fmin <= f < min(fmax, sr/2) (librosa.piptrack freq_mask)."""
    f = np.fft.rfftfreq(N_FFT, 1.0 / sr)
    idx = np.nonzero((max(fmin, 0) <= f) & (f < min(fmax, sr / 2.0)))[0]
    return int(idx[0]), int(idx[-1])


def mel_sparse_layout(bank: np.ndarray):
    """Transposed, slot-padded sparse layout of the mel bank for lane-parallel gathers.

    Filter m is owned by lane m % 32 in slot m // 32; slot s is padded to the longest support in the
    slot so that weight reads are lane-consecutive: melw[(off[s] + i) * 32 + lane].
    """
    lo = np.zeros(N_MELS, dtype=np.int32)
    cnt = np.zeros(N_MELS, dtype=np.int32)
    for m in range(N_MELS):
        nz = np.nonzero(bank[m])[0]
        if nz.size:
            lo[m], cnt[m] = nz[0], nz[-1] - nz[0] + 1
    slot_len = np.array([cnt[32 * s:32 * s + 32].max() for s in range(4)], dtype=np.int32)
    slot_off = np.concatenate(([0], np.cumsum(slot_len)[:-1])).astype(np.int32)
    rows = int(slot_len.sum())
    w = np.zeros((rows, 32), dtype=np.float32)
    for m in range(N_MELS):
        s, lane = divmod(m, 32)
        for i in range(cnt[m]):
            w[slot_off[s] + i, lane] = bank[m, lo[m] + i]
    return w, lo, slot_off, slot_len


def mel_chunk_layout(sr: int, bank: np.ndarray):
    """Layout for the kernel's contiguous-chunk mel projection (lane l owns bins [32l, 32l+32), lane 31 also 1024).

    Every rFFT bin lies in one interval [edge_i, edge_{i+1}) of the 130 Slaney edges, i.e. on the falling slope
    of filter i-1 (weight wa) and the rising slope of filter i (weight wb).  A lane walks its bins with two running
    sums and flushes one partial sum whenever the interval index advances (bit j of mel_mask[lane]); the partial of
    filter m written by lane L sits at slot L*mel_ps + (m - i0[L] + 1).  mel_src[m] lists the (<= 3) slots to add up.
    Requires the interval index to advance by at most 1 per bin (true for sr <= ~60 kHz); raises otherwise.
    """
    fftfreqs = np.fft.rfftfreq(N_FFT, 1.0 / sr)
    edges_hz = _mel_to_hz(np.linspace(_hz_to_mel(0.0), _hz_to_mel(sr / 2.0), N_MELS + 2))
    iv = np.clip(np.searchsorted(edges_hz, fftfreqs, side="right") - 1, 0, N_MELS)
    if not bank[:, N_BINS - 1].any():
        iv[N_BINS - 1] = iv[N_BINS - 2]          # common case: no weight on the Nyquist bin, no flush needed
    if np.diff(iv).max() > 1:
        raise ValueError(f"sample rate {sr}: mel intervals narrower than an FFT bin are not supported")
    wa = np.zeros(N_BINS, dtype=np.float32)
    wb = np.zeros(N_BINS, dtype=np.float32)
    for k in range(N_BINS):
        i = iv[k]
        if 1 <= i <= N_MELS:
            wa[k] = bank[i - 1, k]
        if i <= N_MELS - 1:
            wb[k] = bank[i, k]
    recon = np.zeros_like(bank)
    for k in range(N_BINS):
        i = iv[k]
        if 1 <= i <= N_MELS:
            recon[i - 1, k] += wa[k]
        if i <= N_MELS - 1:
            recon[i, k] += wb[k]
    if not np.array_equal(recon, bank):
        raise ValueError("mel bank is not expressible as adjacent falling/rising slopes")
    ab = np.zeros((33, 32, 2), dtype=np.float32)
    mask = np.zeros(32, dtype=np.uint32)
    i0 = np.zeros(32, dtype=np.int32)
    flush32 = int(iv[N_BINS - 1] != iv[N_BINS - 2])
    touched = []
    for lane in range(32):
        i0[lane] = iv[32 * lane]
        tl = set()
        for j in range(33 if lane == 31 else 32):
            k = 32 * lane + j
            ab[j, lane] = (wa[k], wb[k])
            if 0 < j < 32 and iv[k] != iv[k - 1]:
                mask[lane] |= np.uint32(1) << np.uint32(j)
            tl.update((iv[k] - 1, iv[k]))
        touched.append(tl)
    nflush = [bin(int(mask[l])).count("1") + 2 + (flush32 if l == 31 else 0) for l in range(32)]
    ps = max(nflush) | 1                       # odd row stride of the partial-sum slots
    if 1088 + 32 * ps + 1 > 2112:
        raise ValueError(f"sample rate {sr}: too many mel filters per 32-bin chunk for the kernel's tile")
    zero_slot = 32 * ps
    src = np.full((N_MELS, 3), zero_slot, dtype=np.int32)
    nsrc = np.zeros(N_MELS, dtype=np.int32)
    for lane in range(32):
        for m in sorted(touched[lane]):
            if 0 <= m < N_MELS:
                q = m - (i0[lane] - 1)
                assert 0 <= q < nflush[lane] and nsrc[m] < 3
                src[m, nsrc[m]] = lane * ps + q
                nsrc[m] += 1
    return dict(mel_ab=ab, mel_mask=mask, mel_i0=i0, mel_src=src, mel_flush32=flush32, mel_ps=int(ps))


@functools.lru_cache(maxsize=4)
def build_tables(sr: int = 22050) -> dict:
    n = np.arange(N_FFT, dtype=np.float64)
    hann = (0.5 - 0.5 * np.cos(2.0 * np.pi * n / N_FFT)).astype(np.float32)
    lane = np.arange(32, dtype=np.float64)[None, :]
    k = np.arange(32, dtype=np.float64)[:, None]
    a1 = 2.0 * np.pi * lane * k / 1024.0
    tw1 = np.stack([np.cos(a1), -np.sin(a1)], axis=-1).astype(np.float32)
    a2 = 2.0 * np.pi * (lane + 32.0 * k) / 2048.0
    tw2 = np.stack([0.5 * np.cos(a2), 0.5 * np.sin(a2)], axis=-1).astype(np.float32)
    mb = mel_bank(sr)
    melw, mel_lo, mel_off, mel_len = mel_sparse_layout(mb)
    edges = tuning_edges()
    chroma = np.zeros((N_TUNINGS, N_CHROMA, P_STRIDE), dtype=np.float32)
    for i in range(N_TUNINGS):
        chroma[i, :, :N_BINS] = chroma_bank(sr, float(edges[i]))
    kmin, kmax = piptrack_bin_range(sr)
    chunk = mel_chunk_layout(sr, mb)
    c_hi, c_lo = split_fp16(chroma)
    chroma16 = np.ascontiguousarray(np.stack([c_hi, c_lo], axis=1))          # [100][2][12][1056] float16
    chroma_ny = np.ascontiguousarray(chroma[:, :, N_BINS - 1])               # [100][12] float32 (Nyquist bin)
    chroma_frag = chroma_fragments(c_hi, c_lo)                               # [100][32][2][2][32][4] uint32
    chroma_umma = chroma_umma_images(c_hi, c_lo)                             # [100][16][4096] uint8
    return dict(**chunk, sr=sr, hann=hann, tw1=np.ascontiguousarray(tw1), tw2=np.ascontiguousarray(tw2),
                mel_dense=mb, melw=melw, mel_lo=mel_lo, mel_off=mel_off, mel_len=mel_len,
                chroma16=chroma16, chroma_ny=chroma_ny, chroma_frag=chroma_frag, chroma_umma=chroma_umma, chroma_f32=chroma, dct=dct_matrix(), edges=edges, pip_kmin=kmin, pip_kmax=kmax)
