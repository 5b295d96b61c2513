"""numpy model of the kernel's per-warp 2048-point real FFT (lane/register structure, same tables).

Mirrors the frame loop of csrc/sfx_kernels.cu: z[m] = x[2m] + i x[2m+1]; m = 32*m1 + lane;
stage 1: radix-2 FFT-32 over m1 in registers, outputs in bit-reversed slots (the kernel uses the DIT form on the
bit-reversed register view, this model the equivalent DIF form), twiddle tw1[k1][lane],
exchange through a [32][33] tile, stage 2: FFT-32 over m2, then the real-FFT unpack with the
partner bin held by lane (32-lane)&31, register 31-k2 (lane 0: register (32-k2)&31).
"""
import numpy as np


def brev5(k):
    return int(f"{k:05b}"[::-1], 2)


def dif_fft32(v):
    """v: complex array [32, ...] -> in-place radix-2 DIF, result in bit-reversed order."""
    v = v.copy()
    h = 16
    while h >= 1:
        for b in range(0, 32, 2 * h):
            for j in range(h):
                a, c = v[b + j].copy(), v[b + j + h].copy()
                v[b + j] = a + c
                v[b + j + h] = (a - c) * np.exp(-2j * np.pi * (j * (16 // h)) / 32)
        h //= 2
    return v


def warp_rfft2048(xw, tw1, tw2):
    """xw: windowed frame float[2048]; tw1/tw2: tables float32[32][32][2] -> X[1025] complex."""
    z = xw[0::2] + 1j * xw[1::2]                      # [1024]
    reg = z.reshape(32, 32)                            # reg[m1][lane]
    A = dif_fft32(reg.astype(np.complex128))           # A[brev(k1)][lane]
    ex = np.zeros((32, 33), dtype=np.complex128)
    for k1 in range(32):
        t = tw1[k1, :, 0] + 1j * tw1[k1, :, 1]
        ex[k1, :32] = A[brev5(k1)] * t                 # lane = m2 writes ex[k1][lane]
    b = np.zeros((32, 32), dtype=np.complex128)        # b[m2][lane=k1]
    for m2 in range(32):
        b[m2] = ex[:, m2]                              # lane k1 reads ex[lane][m2]
    C = dif_fft32(b)                                   # C[brev(k2)][lane] = Z[lane + 32*k2]
    X = np.zeros(1025, dtype=np.complex128)
    lanes = np.arange(32)
    for k2 in range(32):
        zv = C[brev5(k2)]
        shf = C[brev5(31 - k2)][(32 - lanes) & 31]     # __shfl_sync(v[brev(31-k2)], (32-lane)&31)
        own = C[brev5((32 - k2) & 31)]
        p = np.where(lanes == 0, own, shf)
        er, ei = zv.real + p.real, zv.imag - p.imag
        orr, oi = zv.real - p.real, zv.imag + p.imag
        c, s = tw2[k2, :, 0], tw2[k2, :, 1]
        xr = 0.5 * er + (c * oi - s * orr)
        xi = 0.5 * ei - (c * orr + s * oi)
        X[lanes + 32 * k2] = xr + 1j * xi
    z0 = C[brev5(0)][0]
    X[1024] = z0.real - z0.imag
    return X
