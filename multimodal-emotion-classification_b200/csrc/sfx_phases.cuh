// sfx_phases.cuh -- the two phase functions shared by the fused kernel (sfx_kernels.cu) and the split kernels
// (sfx_split.cu):
//   process_frame<kDebug, kSplit> : phase 1 for one STFT frame, executed by one warp
//   clip_tail<kDebug>             : phases 2-3 + the pooled output row for one clip, executed by one CTA
// kSplit selects where per-frame results are kept: the fused kernel accumulates centroid / roll-off / zero crossings /
// log-mel max per warp in shared memory and appends peak records to the warp's own segment of the clip's record buffer
// (fill level in a register); the split pipeline stores the per-frame values in the clip's slice (summed later in a
// fixed order) and reserves record space with one global atomic per frame.
#pragma once
#include "sfx_device.cuh"

// Which dead scratch the fused tail drops from L2 (discard_l2_range: discard.global.L2, SASS CCTL.E.RML2, one per 128-byte line):
// 1 peak records (after the per-peak loop), 2 log-mel rows (after the pooling), 4 FP16 |X|^2 rows (after the chroma projection).
// Measured on the bench mix (tools/dram_ab.sh, tools/ab_probe.py): 4 takes the DRAM traffic from 549 to 485 KB per clip (7 takes
// it to 450) but costs 1.3 % of the throughput (2 145 cache-control instructions per clip through the LSU), 1 and 2 save ~10 KB
// for 1 %: HBM is at 12 % utilisation, so the default is 0 (off).
#ifndef SFX_DISCARD
#define SFX_DISCARD 0
#endif

namespace sfx {

// Cycle accounting of the fused kernel (library built with -DSFX_FUSED_DIAG, tools/fused_prof.py): thread 0 of every CTA adds
// the cycles between phase boundaries to g_fprof[phase]; 0 frames, 1 per-peak loop, 2 median select + histogram, 3 MFCC,
// 4 wait for the bank, 5 chroma, 6 epilogue, 7 clips
#ifdef SFX_FUSED_DIAG
__device__ unsigned long long g_fprof[16];   // 8..10 = finer marks inside phase 2 (before / after the radix select, upper median)
#define FPROF_MARK(k) do { if (threadIdx.x == 0) { const long long fp1 = clock64(); atomicAdd(&g_fprof[k], static_cast<unsigned long long>(fp1 - fprof_t)); fprof_t = fp1; } } while (0)
#else
#define FPROF_MARK(k) do { } while (0)
#endif

// shared-memory tables and the calling warp's tile (phase 1)
struct FrameSmem {
    const float2* s_hann;    // 32 x (cos, cos', sin, sin'): the Hann phases of a lane's two samples per row (fill_hann_phases)
    const float2* s_tw1; const float2* s_tw2; const float2* s_melab;
    float* Pb;            // this warp's exchange / |X|^2 tile [kExFloats]
    float2* ex;           // the same tile as float2
    float* part;          // mel partial sums inside the tile
    unsigned mel_mask; int mel_ps; int msrc[4];
    const int* s_msrc;    // fused kernel: the msrc words of all 128 filters in shared memory (read per frame; msrc[] unused)
    float bin_hz; bool aligned8;
};

// where phase 1 leaves its results
struct FrameOut {
    __half* gP16; float* gL; float4* gRec; float* gE; float* gNy; float* gInvS;     // the clip's scratch slice
    int* npk;                                                                          // split: the clip's peak counter
    float4* gSeg;                                                                      // fused: this warp's record segment
    double* s_wacc; float* s_f;                                                        // fused: per-warp accumulators
    double* s_lm;         // fused: this warp's float64 sums of the log-mel rows of its non-zero frames [128]
    float* s_lmin;        // fused: this warp's per-lane minimum of those rows [32]
    float* gCent; float* gRoll; float* gLmax; int* gZc;                                // split: per-frame values
    float* gFv;           // stream: per-frame values as one record of kFvStride floats per frame (one live pointer
                          // instead of seven): [0] scaled Nyquist |X|^2, [1] 1/scale, [2] hop energy, [3] centroid,
                          // [4] roll-off, [5] row max of log-mel, [6] weighted zero crossings (int), [7] unused
    int* cursor;          // stream: fill level of the slot's dense peak-record array gRec (shared memory, atomicAdd)
};
constexpr int kFvStride = 8;

// shared memory of phases 2-3
struct ClipSmem {
    float* s_ex; double* s_pool; double* s_wacc; double* s_edges; unsigned long long* s_mbar; int* s_hist; int* s_i; float* s_f;
    const double* s_lm;   // fused: [kWarps][128] per-warp float64 sums of the unclamped log-mel rows of the non-zero frames,
    const float* s_lmin;  //        [kWarps][32] per-lane minima of those rows, s_f[16 + w] = all-zero frames seen by warp w
                          //        (nullptr in the split pipeline: phase 3a then always reads the rows back)
    unsigned long long* s_ubar;   // UMMA build of the fused kernel: 7 mbarriers = full[3], empty[3], accumulator ready
    unsigned tmem;                //   base address of the CTA's 32 tensor-memory columns
};

// running state of the UMMA chroma pipeline of one CTA (registers, identical in every thread): operand blocks issued since
// the kernel started (stage = blocks % 3, mbarrier parity = (blocks / 3) & 1) and accumulator tiles completed
struct UmmaState { unsigned blocks = 0, tiles = 0; };

struct ClipSlice {
    __half* gP16; float* gL; float4* gRec; unsigned* gKey; float* gE; float* gNy; float* gInvS; unsigned char* gBin;
    int seg_cap;          // peak records are kept in kWarps segments of this capacity (cs.s_i[20 + w] = records in segment w)
};

// ------------------------------------------------------------------------------------------------ phase 1
// kMode: where per-frame results are kept.
//   kModeFused : per-warp accumulators in shared memory, peaks appended to the warp's own record segment
//   kModeSplit : per-frame values in the clip's slice, record space reserved with one global atomic per frame
//   kModeStream: per-frame values in the clip's slice (summed in frame order by the tail warp: deterministic whatever
//                warp ran the frame), peaks appended to the calling warp's own record segment of the clip's slot
constexpr int kModeFused = 0, kModeSplit = 1, kModeStream = 2, kModeFusedUmma = 3;   // 3 = fused, |X|^2 rows stored as tcgen05 operand images

template <bool kDebug, int kMode>
__device__ __forceinline__ void process_frame(const Params& p, const DevTables& tb, const FrameSmem& fs, const FrameOut& fo,
                                              const float* __restrict__ x, const long long n, const int T, const int t,
                                              const int clip, const int lane, const int warp, int& acc_zc, int& wcount) {
    constexpr bool kSplit = kMode == kModeSplit;          // records through a global atomic per frame
    constexpr bool kFused = kMode == kModeFused || kMode == kModeFusedUmma;
    constexpr bool kUmmaRows = kMode == kModeFusedUmma;    // |X|^2 rows go into the shared-memory images of the UMMA A operand
    constexpr bool kFrameVals = !kFused;                   // per-frame centroid / roll-off / zero crossings / row max stored
    constexpr bool kStream = kMode == kModeStream;         // ... as one record per frame (fo.gFv)
    float re[32], im[32];
    load_frame(x, n, t, lane, fs.aligned8, re, im);
    // ---- all-zero frame (the zero tail load_audio pads short clips with, reference :15-16): every result is
    //      known in closed form and equals what the general path computes from zeros, so the FFT is skipped:
    //      |X|^2 = 0, log-mel = 10*log10(1e-10) in all bands, no peaks, centroid = rolloff = 0, hop energy 0,
    //      no sign changes.
    {
        unsigned any_bits = 0;
#pragma unroll
        for (int m1 = 0; m1 < 32; ++m1) any_bits |= __float_as_uint(re[m1]) | __float_as_uint(im[m1]);
        if (!__any_sync(0xffffffffu, (any_bits & 0x7fffffffu) != 0u)) {
            const float lm0 = 3.01029995663981195f * __log2f(1e-10f);
            float* Lg = fo.gL + static_cast<size_t>(t) * kMels;
#pragma unroll
            for (int s4 = 0; s4 < 4; ++s4) Lg[32 * s4 + lane] = lm0;
            if constexpr (kUmmaRows) {
                unsigned char* img = reinterpret_cast<unsigned char*>(fo.gP16);
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4)
                    *reinterpret_cast<uint4*>(img + umma_p16_offset(p.Tmax, t, lane >> 1, (lane & 1) * 4 + q4)) = make_uint4(0u, 0u, 0u, 0u);
            } else {
                uint4* dst = reinterpret_cast<uint4*>(fo.gP16 + static_cast<size_t>(t) * kP16Row + 32 * lane);
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) dst[q4] = make_uint4(0u, 0u, 0u, 0u);
            }
            if constexpr (kStream) {
                if (lane < kFvStride) fo.gFv[static_cast<size_t>(t) * kFvStride + lane] = lane == 5 ? lm0 : 0.0f;
            } else if (lane == 0) {
                fo.gE[t] = 0.0f;
                fo.gNy[t] = 0.0f;
                // split pipeline: -1 marks the all-zero frame for the clips kernel (a non-zero frame below 2^-113 also has
                // 1/scale = 0); where the chroma normalisation meets it, 0 * -1 = -0 adds nothing, like the fused kernel's 0
                fo.gInvS[t] = kFrameVals ? -1.0f : 0.0f;
                if constexpr (kFrameVals) { fo.gCent[t] = 0.0f; fo.gRoll[t] = 0.0f; fo.gLmax[t] = lm0; fo.gZc[t] = 0; }
                else { fo.s_f[warp] = fmaxf(fo.s_f[warp], lm0); fo.s_f[16 + warp] += 1.0f; }
            }
            if (kDebug) {
                if (t < p.dbg.T_dbg) {
                    if (p.dbg.P) {
                        float* dP = p.dbg.P + (static_cast<size_t>(clip) * p.dbg.T_dbg + t) * kPStride;
                        for (int k = lane; k < kBins; k += 32) dP[k] = 0.0f;
                    }
                    if (p.dbg.logmel)
                        for (int m = lane; m < kMels; m += 32)
                            p.dbg.logmel[(static_cast<size_t>(clip) * p.dbg.T_dbg + t) * kMels + m] = lm0;
                    if (p.dbg.frame_feat && lane == 0) {
                        float* ff = p.dbg.frame_feat + (static_cast<size_t>(clip) * p.dbg.T_dbg + t) * 4;
                        ff[0] = 0.f; ff[1] = 0.f; ff[3] = 0.f;
                    }
                }
            }
            __syncwarp();
            return;
        }
    }
    if constexpr (kFused) {
        if (lane == 0) fo.s_f[8 + warp] = static_cast<float>(t);       // this warp's last non-zero frame so far (t grows)
    }
    // ---- energy of hop t (samples [512t, 512t+512) = rows 16..23); librosa.feature.rms of frame t is
    //      sqrt((E[t-2] + E[t-1] + E[t] + E[t+1]) / 2048) and is pooled in the epilogue
    {
        float he = 0.0f;
#pragma unroll
        for (int m1 = 16; m1 < 24; ++m1) { he = fmaf(re[m1], re[m1], he); he = fmaf(im[m1], im[m1], he); }
        he = warp_sum(he);
        if (lane == 0) {
            if constexpr (kStream) fo.gFv[static_cast<size_t>(t) * kFvStride + 2] = he;
            else fo.gE[t] = he;
        }
    }

    // ---- zero crossings of hop t (samples [512t, 512t+512)), weighted by how many frames see them
    int zc_hop = 0, zc_w = 0;
    {
        // librosa.zero_crossings(threshold=1e-10, zero_pos=True): sign(x) := (double)x < -1e-10, which for float32
        // x is exactly x < -9.99999944e-11f (0xaedbe6fe, the smallest float32 not below -1e-10)
        const float zthr = __uint_as_float(0xaedbe6feu);
        // sign bits of the hop's samples: SA bit r = sample 512t + 64r + 2*lane, SB bit r+1 = the sample after it,
        // SB bit 0 = this lane's last sample of the previous row (row 15 = the 64 samples before the hop)
        unsigned SA = 0u, SB = (im[15] < zthr) ? 1u : 0u;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            SA |= (re[16 + r] < zthr) ? (1u << r) : 0u;
            SB |= (im[16 + r] < zthr) ? (2u << r) : 0u;
        }
        const unsigned up = __shfl_up_sync(0xffffffffu, SB, 1);
        const unsigned wrap = __shfl_sync(0xffffffffu, SB, 31);
        const unsigned PB = (lane == 0) ? wrap : (up >> 1);          // bit r: sign of the sample before row r's first
        const int ib = kHop * t + 2 * lane;
        const int nm1 = static_cast<int>(n) - 1;
        // crossings are counted at positions 1 <= i <= n-1: rows r with ib + 64 r (+1) <= n-1
        const int v0 = min(max((nm1 - ib + 64) >> 6, 0), 8), v1 = min(max((nm1 - ib + 63) >> 6, 0), 8);
        unsigned m0 = (1u << v0) - 1u;
        if (ib == 0) m0 &= ~1u;
        const unsigned c0 = (SA ^ PB) & m0, c1 = (SA ^ (SB >> 1)) & ((1u << v1) - 1u);
        const int zc = __popc(c0) + __popc(c1);
        const int zc_first = (lane == 0) ? static_cast<int>(c0 & 1u) : 0;
        const int tlo = max(0, t - 1);
        const int multA = min(T - 1, t + 2) - tlo + 1;
        const int mult0 = min(T - 1, t + 1) - tlo + 1;
        zc_hop = zc;
        zc_w = multA * zc - (multA - mult0) * zc_first;
    }

    // ---- Hann window (periodic, scipy.signal.get_window('hann', 2048)) fused into the first radix-2 stage of the
    //      radix-32 pass over m1.  From here on each complex sample z[m1] = (x[2n], x[2n+1]), n = lane + 32 m1, is one
    //      packed register pair; the window table holds rows r and r + 16 of a lane side by side.
    c64 z[32];
    {
        // No window table: sample n = 64 r + 2 lane + e of the frame has phase theta(lane, e) + r * 2 pi / 32, so
        //   w[n] = 0.5 - 0.5 cos(theta + r phi) = 0.5 - (0.5 cos r phi) cos theta + (0.5 sin r phi) sin theta
        // is two packed FMAs on the lane's (cos theta, sin theta) pairs (one 128-bit load per frame) with immediate row
        // constants, and row r + 16 is half a turn further: w[n + 1024] = 1 - w[n].  Saves the 64 shared-memory wavefronts
        // per frame of an 8 KB table on the kernel's busiest pipe; values within 1 ulp(0.5) of the rounded exact ones.
        const float4 cs4 = reinterpret_cast<const float4*>(fs.s_hann)[lane];
        const c64 C = pk(cs4.x, cs4.y), S = pk(cs4.z, cs4.w);
        sfor<16>([&](auto R) {
            constexpr int r = decltype(R)::value;
            const c64 w = fma2(C, bc2(-0.5f * kCos32[r]), fma2(S, bc2(0.5f * kSin32[r]), bc2(0.5f)));
            const c64 w16 = sub2(bc2(1.0f), w);
            const c64 za = mul2(pk(re[r], im[r]), w);
            const c64 xb = pk(re[r + 16], im[r + 16]);
            z[r] = fma2(xb, w16, za);
            z[r + 16] = fma2(xb, neg2(w16), za);
        });
    }
    // ---- 1024-pt complex FFT: radix-32 over m1, twiddle, transpose, radix-32 over m2
    fft32p<2>(z);
    // exchange tile: float4 slot (k1/2)*33 + lane holds rows k1, k1+1 of column `lane`
    sfor<16>([&](auto K) {
        constexpr int k1 = 2 * decltype(K)::value;
        constexpr int a = brev5(k1), b = brev5(k1 + 1);
        const float4 w = *reinterpret_cast<const float4*>(fs.s_tw1 + (k1 >> 1) * 64 + 2 * lane);
        ulonglong2 v;
        v.x = cmul(z[a], w.x, w.y);
        v.y = cmul(z[b], w.z, w.w);
        reinterpret_cast<ulonglong2*>(fs.ex)[(k1 >> 1) * 33 + lane] = v;
    });
    __syncwarp();
    {
        const c64* src = reinterpret_cast<const c64*>(fs.ex) + ((lane >> 1) * 33) * 2 + (lane & 1);   // row `lane`
#pragma unroll
        for (int m2 = 0; m2 < 32; ++m2) z[m2] = src[2 * m2];
    }
    __syncwarp();
    fft32p(z);

    // ---- real-FFT unpack, one conjugate pair per step: lane holds Z[lane + 32*k2]; for k2 < 16 it forms
    //      X[k] and X[1024-k] (k = lane + 32*k2) from Z[k] and Z[1024-k] (lane (32-lane)&31, register 31-k2;
    //      lane 0 pairs with its own register (32-k2)&31).  Bin 512 (self-paired) is lane 0's register 16.
    //      With E = Z[k] + conj(Z[1024-k]), O = Z[k] - conj(Z[1024-k]) and w = exp(-2 pi i k / 2048):
    //      X[k] = E/2 - i w O/2 and conj(X[1024-k]) = E/2 + i w O/2; only |.|^2 is kept.
    float pmax = 0.0f;
    {
        float* PbL = fs.Pb + lane;
        float* PbU = fs.Pb + (lane == 0 ? 1152 : 1148 - lane);
        const int plane = (32 - lane) & 31;
        float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
        sfor<16>([&](auto K) {
            constexpr int k2 = decltype(K)::value;
            constexpr int a = brev5(k2), b = brev5(31 - k2), c = brev5((32 - k2) & 31);
            float br_, bi_, cr_, ci_;
            upk(z[b], br_, bi_);
            upk(z[c], cr_, ci_);
            float pr = __shfl_sync(0xffffffffu, br_, plane);
            float pi = __shfl_sync(0xffffffffu, bi_, plane);
            if (lane == 0) { pr = cr_; pi = ci_; }
            if constexpr ((k2 & 1) == 0) w4 = *reinterpret_cast<const float4*>(fs.s_tw2 + (k2 >> 1) * 64 + 2 * lane);
            const float wc = (k2 & 1) ? w4.z : w4.x, ws = (k2 & 1) ? w4.w : w4.y;   // (0.5 cos, 0.5 sin)(2 pi k / 2048)
            const c64 pc = pk(pr, -pi);                                  // conj(Z[1024-k])
            const c64 E = add2(z[a], pc), O = sub2(z[a], pc);
            const c64 q = fma2(rot_mi(O), bc2(wc), mul2(O, bc2(-ws)));   // -i w O / 2
            const c64 X = fma2(E, bc2(0.5f), q), Y = fma2(E, bc2(0.5f), neg2(q));
            float xr, xi, yr, yi;
            upk(X, xr, xi);
            upk(Y, yr, yi);
            const float P = fmaf(xr, xr, xi * xi);
            const float Q = fmaf(yr, yr, yi * yi);
            pmax = fmaxf(pmax, fmaxf(P, Q));
            PbL[kPRow * k2] = P;
            PbU[-kPRow * k2] = Q;
        });
        if (lane == 0) {
            constexpr int h = brev5(16);
            float hr, hi;
            upk(z[h], hr, hi);
            const float P512 = fmaf(hr, hr, hi * hi);
            fs.Pb[pidx(512)] = P512;
            pmax = fmaxf(pmax, P512);
        }
    }
    pmax = warp_max(pmax);
    __syncwarp();
    if (kDebug) {
        if (p.dbg.P && t < p.dbg.T_dbg) {
            float* dP = p.dbg.P + (static_cast<size_t>(clip) * p.dbg.T_dbg + t) * kPStride;
            for (int k = lane; k < kBins; k += 32) dP[k] = fs.Pb[pidx(k)];
        }
    }

    // ---- contiguous pass: lane owns bins [32*lane, 32*lane+32) (+1024 for lane 31):
    //      |X| prefix sums (centroid, 0.85 roll-off) and the Slaney mel projection as running
    //      falling/rising partial sums flushed whenever the filter interval advances
    float cent_t, roll_t;
    {
        float run = 0.0f, ks = 0.0f;
        c64 acc = pk(0.0f, 0.0f);                        // (falling, rising) partial sums of the current interval
        float* pq = fs.part + lane * fs.mel_ps;
        const float4* Prow = reinterpret_cast<const float4*>(fs.Pb + kPRow * lane);
        // the chroma phase reads this frame's |X|^2 back as FP16 scaled by an exact power of two that puts the
        // frame maximum in [2^14, 2^15) (the per-frame inf-norm of chroma_stft cancels the scale)
        // biased exponent of the scale = 14 - (E - 127) + 127 = 268 - E, clamped to a finite power of two
        const unsigned sbits = min(268u - ((__float_as_uint(pmax) >> 23) & 255u), 254u) << 23;
        const float scale = __uint_as_float(sbits);
        unsigned h2[16];
        ulonglong2 ab = make_ulonglong2(0ull, 0ull);
        sfor<8>([&](auto G) {
            constexpr int g = decltype(G)::value;
            const float4 P4 = Prow[g];
            {
                float a, b;
                upk(mul2(pk(P4.x, P4.y), bc2(scale)), a, b);
                h2[2 * g] = pack_half2(a, b);
                upk(mul2(pk(P4.z, P4.w), bc2(scale)), a, b);
                h2[2 * g + 1] = pack_half2(a, b);
            }
            sfor<4>([&](auto Q4) {
                constexpr int q = decltype(Q4)::value, j = 4 * g + q;
                const float P = q == 0 ? P4.x : q == 1 ? P4.y : q == 2 ? P4.z : P4.w;
                if (j > 0 && ((fs.mel_mask >> j) & 1u)) {
                    float lo, hi;
                    upk(acc, lo, hi);
                    *pq++ = lo;
                    acc = pk(hi, 0.0f);
                }
                if constexpr ((j & 1) == 0) ab = *reinterpret_cast<const ulonglong2*>(fs.s_melab + (j >> 1) * 64 + 2 * lane);
                acc = fma2(bc2(P), (j & 1) ? ab.y : ab.x, acc);
                const float sv = sqrt_approx(P);
                run += sv;
                ks = fmaf(static_cast<float>(j), sv, ks);
            });
        });
        float accA, accB;
        upk(acc, accA, accB);
        if (lane == 31) {
            const float P = fs.Pb[pidx(1024)];
            if (tb.mel_flush32) { *pq++ = accA; accA = accB; accB = 0.0f; }
            const float2 ab1 = fs.s_melab[16 * 64 + 2 * 31];
            accA = fmaf(ab1.x, P, accA);
            accB = fmaf(ab1.y, P, accB);
            const float sv = sqrt_approx(P);
            run += sv;
            ks = fmaf(32.0f, sv, ks);
        }
        pq[0] = accA;
        pq[1] = accB;
        if (lane == 0) fs.part[32 * fs.mel_ps] = 0.0f;   // zero slot read by filters with < 3 contributing lanes
        {
            if constexpr (kUmmaRows) {
                // the row goes straight into the shared-memory images of the tcgen05 A operand (K-major, 128-byte swizzle):
                // this lane's 32 bins are 4 of the 8 16-byte chunks of K block lane / 2
                unsigned char* img = reinterpret_cast<unsigned char*>(fo.gP16);
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4)
                    *reinterpret_cast<uint4*>(img + umma_p16_offset(p.Tmax, t, lane >> 1, (lane & 1) * 4 + q4)) =
                        make_uint4(h2[4 * q4], h2[4 * q4 + 1], h2[4 * q4 + 2], h2[4 * q4 + 3]);
            } else {
                uint4* dst = reinterpret_cast<uint4*>(fo.gP16 + static_cast<size_t>(t) * kP16Row + 32 * lane);
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) dst[q4] = make_uint4(h2[4 * q4], h2[4 * q4 + 1], h2[4 * q4 + 2], h2[4 * q4 + 3]);
            }
            if constexpr (kStream) {
                if (lane == 31) fo.gFv[static_cast<size_t>(t) * kFvStride] = fs.Pb[pidx(1024)] * scale;
                if (lane == 0) fo.gFv[static_cast<size_t>(t) * kFvStride + 1] = __uint_as_float((254u << 23) - sbits);
            } else {
                if (lane == 31) fo.gNy[t] = fs.Pb[pidx(1024)] * scale;
                if (lane == 0) fo.gInvS[t] = __uint_as_float((254u << 23) - sbits);      // 1/scale, exact
            }
        }
        float inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        float exc = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) exc = 0.0f;
        const float total = __shfl_sync(0xffffffffu, inc, 31);
        const float thr = __fmul_rn(0.85f, total);
        // first bin whose cumulative |X| reaches the threshold.  The prefix sums are monotone, so the lanes whose whole
        // 32-bin chunk lies below the threshold form a prefix of the warp; the chunk of the first other lane L holds the
        // bin, and the warp finds it there together: lane j re-reads bin 32 L + j from the tile, one shuffle scan, one
        // ballot (instead of every lane keeping its 32 running sums in registers and counting through them)
        const int L = __popc(__ballot_sync(0xffffffffu, inc < thr));
        int first = 1024;
        if (L < 32) {                                                  // warp-uniform
            const float thrL = thr - __shfl_sync(0xffffffffu, exc, L);
            float pre = sqrt_approx(fs.Pb[kPRow * L + lane]);
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float v = __shfl_up_sync(0xffffffffu, pre, o);
                if (lane >= o) pre += v;
            }
            first = min(1024, 32 * L + __popc(__ballot_sync(0xffffffffu, pre < thrL)));
        }
        const float num = warp_sum(fmaf(32.0f * lane, run, ks));
        cent_t = (total < FLT_MIN) ? 0.0f : (num / total) * fs.bin_hz;
        roll_t = static_cast<float>(first) * fs.bin_hz;
    }
    if constexpr (kStream) {
        const int zsum = warp_sum_i(zc_w);
        if (lane == 0) {
            float* fv = fo.gFv + static_cast<size_t>(t) * kFvStride;
            fv[3] = cent_t; fv[4] = roll_t; fv[6] = __int_as_float(zsum);
        }
    } else if constexpr (kFrameVals) {
        const int zsum = warp_sum_i(zc_w);
        if (lane == 0) { fo.gCent[t] = cent_t; fo.gRoll[t] = roll_t; fo.gZc[t] = zsum; }
    } else {
        acc_zc += zc_w;
        if (lane == 0) {
            fo.s_wacc[warp * 16 + 0] += static_cast<double>(cent_t);
            fo.s_wacc[warp * 16 + 1] += static_cast<double>(roll_t);
        }
    }
    __syncwarp();

    // ---- log-mel rows: filter m = 32*s + lane adds its (<= 3) partial sums in a fixed order
    {
        float* Lg = fo.gL + static_cast<size_t>(t) * kMels;
        float gmax = -FLT_MAX, gmin = FLT_MAX;
        // every load of the section is issued before its first store (the compiler cannot tell that the running sums and
        // the partial-sum slots do not overlap, and would otherwise chain each band's loads behind the previous band's
        // store): slot words, partial sums, then the running float64 sums
        int ms4[4];
        float pa[4], pb[4], pc[4];
        double lsum[4];
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
            if constexpr (kFused) ms4[s4] = fs.s_msrc[32 * s4 + lane];
            else ms4[s4] = fs.msrc[s4];
        }
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
            pa[s4] = fs.part[ms4[s4] & 1023];
            pb[s4] = fs.part[(ms4[s4] >> 10) & 1023];
            pc[s4] = fs.part[ms4[s4] >> 20];
            if constexpr (kFused) lsum[s4] = fo.s_lm[32 * s4 + lane];
        }
        float lmin_old = 0.0f;
        if constexpr (kFused) lmin_old = fo.s_lmin[lane];
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
            const float mel = (pa[s4] + pb[s4]) + pc[s4];
            const float lm = 3.01029995663981195f * __log2f(fmaxf(1e-10f, mel));   // 10*log10(x)
            Lg[32 * s4 + lane] = lm;
            gmax = fmaxf(gmax, lm);
            if constexpr (kFused) {       // running sums / minimum for the clamp-free MFCC pooling (phase 3a)
                fo.s_lm[32 * s4 + lane] = lsum[s4] + static_cast<double>(lm);
                gmin = fminf(gmin, lm);
            }
            if (kDebug) {
                if (p.dbg.logmel && t < p.dbg.T_dbg)
                    p.dbg.logmel[(static_cast<size_t>(clip) * p.dbg.T_dbg + t) * kMels + 32 * s4 + lane] = lm;
            }
        }
        if constexpr (kFused) fo.s_lmin[lane] = fminf(lmin_old, gmin);
        (void)gmin; (void)lsum; (void)lmin_old;
        gmax = warp_max(gmax);
        if (lane == 0) {
            if constexpr (kStream) fo.gFv[static_cast<size_t>(t) * kFvStride + 5] = gmax;
            else if constexpr (kFrameVals) fo.gLmax[t] = gmax;
            else fo.s_f[warp] = fmaxf(fo.s_f[warp], gmax);
        }
    }

    // ---- piptrack peak detection on the power spectrum (bins kmin..kmax); the per-peak arithmetic
    //      (parabolic shift, pitch, tuning residual) is done in phase 2 on the compacted records
    if constexpr (kStream) {
        // one pass per chunk of <= 14 rows: the chunk's peaks are ballot-compacted into the tile's mel partial-sum area
        // (free since the log-mel rows above; 239 records, a row holds <= 16 peaks), then appended, densely and with
        // coalesced 128-bit stores, to the slot's record array behind one shared-memory atomicAdd.  The tail reads the
        // records as one dense array whatever warp produced them, in whatever order (median and histogram do not care).
        const float ref = __fmul_rn(0.1f, pmax);
        const int kfirst = tb.kmin + lane;
        const float* q0 = fs.Pb + pidx(kfirst - 1);
        const float* q1 = fs.Pb + pidx(kfirst);
        const float* q2 = fs.Pb + pidx(kfirst + 1);
        const int nrows = (tb.kmax - tb.kmin + 32) >> 5;
        const int rmax = (tb.kmax - kfirst) >> 5;
        const unsigned lt = (1u << lane) - 1u;
        float4* stage = reinterpret_cast<float4*>(fs.part);
        constexpr int kChunkRows = (kExFloats - kPartOff) / 4 / 16;          // 14
        for (int r0 = 0; r0 < nrows; r0 += kChunkRows) {
            const int r1 = min(nrows, r0 + kChunkRows);
            unsigned cnt = 0;
#pragma unroll 4
            for (int r = r0; r < r1; ++r) {
                const int off = kPRow * r;
                const float pm = q0[off], pc = q1[off], pp = q2[off];
                const bool pk = (r <= rmax) & (pc > ref) & (pc > pm) & (pc >= pp);
                const unsigned bal = __ballot_sync(0xffffffffu, pk);
                if (pk) stage[cnt + __popc(bal & lt)] = make_float4(pm, pc, pp, __int_as_float(kfirst + 32 * r));
                cnt += __popc(bal);
            }
            __syncwarp();
            if (cnt) {
                int base = 0;
                if (lane == 0) base = atomicAdd(fo.cursor, static_cast<int>(cnt));
                base = __shfl_sync(0xffffffffu, base, 0);
                for (unsigned i = lane; i < cnt; i += 32) fo.gRec[base + i] = stage[i];
            }
            __syncwarp();
        }
    } else if constexpr (!kSplit) {
        // one pass over groups of 128 bins: a lane takes four consecutive bins with one 128-bit load (conflict-free in the
        // padded tile), their outer neighbours come from the adjacent lanes by shuffle (the group's two edge bins by one
        // extra load), so a bin costs a third of the shared-memory wavefronts of three shifted row reads.  The peaks of
        // each of the four bin positions are appended (ballot-compacted) to this warp's own record segment of the clip,
        // whose fill level `wcount` the warp carries in a register -- no atomics, no second pass.  (The order of a frame's
        // records differs from bin order; median and histogram do not depend on it.)
        const float ref = __fmul_rn(0.1f, pmax);
        const unsigned lt = (1u << lane) - 1u;
        unsigned cnt = static_cast<unsigned>(wcount);
        const int kmin = tb.kmin, kmax = tb.kmax;
        const unsigned krange = static_cast<unsigned>(kmax - kmin);       // bin k is inside iff unsigned(k - kmin) <= krange
        for (int gq = kmin >> 7; gq <= (kmax >> 7); ++gq) {
            const int b0 = 128 * gq + 4 * lane;
            const float4 v = *reinterpret_cast<const float4*>(fs.Pb + pidx(b0));
            float edge = 0.0f;
            if ((lane == 0 && gq > 0) || lane == 31) edge = fs.Pb[pidx(lane == 0 ? 128 * gq - 1 : 128 * gq + 128)];
            float vl = __shfl_up_sync(0xffffffffu, v.w, 1), vr = __shfl_down_sync(0xffffffffu, v.x, 1);
            if (lane == 0) vl = edge;
            if (lane == 31) vr = edge;
            const float w[6] = {vl, v.x, v.y, v.z, v.w, vr};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float pm = w[i], pc = w[i + 1], pp = w[i + 2];
                const int k = b0 + i;
                const bool pk = (static_cast<unsigned>(k - kmin) <= krange) & (pc > ref) & (pc > pm) & (pc >= pp);
                const unsigned bal = __ballot_sync(0xffffffffu, pk);
                if (bal) {                                           // warp-uniform: most of a tonal frame holds no peak
                    st_record_if(pk, fo.gSeg, cnt + __popc(bal & lt), pm, pc, pp, k);
                    cnt += __popc(bal);
                }
            }
        }
        wcount = static_cast<int>(cnt);
    } else {
        {
            const float ref = __fmul_rn(0.1f, pmax);
            const int kfirst = tb.kmin + lane;
            const float* q0 = fs.Pb + pidx(kfirst - 1);
            const int d0 = pidx(kfirst) - pidx(kfirst - 1), d1 = pidx(kfirst + 1) - pidx(kfirst - 1);
            const int nrows = (tb.kmax - tb.kmin + 32) >> 5;          // <= 16 (bins 1..1023, 32 per row)
            unsigned flags = 0;                                      // bit r: this lane's bin of row r is a peak
            {
                const float* q = q0;
                const int rlast = tb.kmax - kfirst;                  // rows with 32*r <= rlast hold a bin of the range
#pragma unroll 4
                for (int r = 0; r < nrows; ++r, q += kPRow) {
                    const float pm = q[0], pc = q[d0], pp = q[d1];
                    const bool pk = (32 * r <= rlast) && pc > ref && pc > pm && pc >= pp;
                    flags |= (pk ? 1u : 0u) << r;
                }
            }
            // lane-wise compaction: exclusive prefix of the per-lane peak counts, one atomic per frame
            const int mine = __popc(flags);
            int inc = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            const int total = __shfl_sync(0xffffffffu, inc, 31);
            if (total) {
                int base = 0;
                if (lane == 0) base = atomicAdd(fo.npk, total);
                base = __shfl_sync(0xffffffffu, base, 0);
                float4* dst = fo.gRec + base + (inc - mine);
                while (flags) {
                    const int r = __ffs(flags) - 1;
                    flags &= flags - 1;
                    const float* q = q0 + kPRow * r;
                    *dst++ = make_float4(q[0], q[d0], q[d1], __int_as_float(kfirst + 32 * r));
                }
            }
        }
    }
    if (kDebug) {
        const int zch = warp_sum_i(zc_hop);
        if (p.dbg.frame_feat && t < p.dbg.T_dbg && lane == 0) {
            float* ff = p.dbg.frame_feat + (static_cast<size_t>(clip) * p.dbg.T_dbg + t) * 4;
            ff[0] = cent_t; ff[1] = roll_t; ff[3] = static_cast<float>(zch);
        }
    }
    (void)zc_hop;
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------ per-peak reference forms
// librosa.piptrack's parabolic shift as numba types it: float32 sums, float64 quotient, float32 result
static __device__ __noinline__ float peak_shift_exact(float pm, float pc, float pp) {
    const float sum = __fadd_rn(pp, pm);
    const float dif = __fsub_rn(pp, pm);
    const double a = static_cast<double>(sum) - 2.0 * static_cast<double>(pc);
    const double b = static_cast<double>(dif) * 0.5;
    return (fabs(b) >= fabs(a)) ? 0.0f : static_cast<float>(-b / a);
}
// librosa.pitch_tuning: float32 mod(12*log2(f/27.5), 1) wrapped to [-0.5, 0.5), located in np.linspace(-0.5, 0.5, 101)
static __device__ __noinline__ int peak_bin_exact(float pitch, const double* s_edges) {
    const float o = log2f(__fdiv_rn(pitch, 27.5f));
    const float v = __fmul_rn(12.0f, o);
    float res = v - floorf(v);
    if (res >= 0.5f) res = res - 1.0f;
    const double rd = static_cast<double>(res);
    int bi = static_cast<int>(floor((rd + 0.5) * 100.0));
    bi = max(0, min(kTunings - 1, bi));
    while (bi > 0 && rd < s_edges[bi]) --bi;
    while (bi < kTunings - 1 && rd >= s_edges[bi + 1]) ++bi;
    return bi;
}

// ------------------------------------------------------------------------------------------------ phases 2-3
// Expects (set up by the caller, followed by __syncthreads): cs.s_i[20 + w] = peak records in segment w, cs.s_i[17] = cs.s_i[18] = 0, cs.s_i[19] = ~0, cs.s_f[w] = per-warp log-mel
// max, cs.s_wacc[w*16 + 0/1] = per-warp centroid / roll-off sums, cs.s_i[8+w] = per-warp weighted zero-crossing counts.
// kWarpSegs (fused kernel): the per-peak loop gives segment w to warp w (every warp filled one in phase 1, all about the same
// size) instead of walking one dense index over the segments: no boundary search per record, warp-contiguous loads.
template <bool kDebug, bool kUmmaTail = false, bool kWarpSegs = false>
__device__ __forceinline__ void clip_tail(const Params& p, const DevTables& tb, const ClipSmem& cs, const ClipSlice& sl,
                                          const int clip, const int T, float* __restrict__ out, unsigned& bank_parity,
                                          const int tid, const int lane, const int warp, UmmaState* us = nullptr) {
#ifdef SFX_FUSED_DIAG
    long long fprof_t = clock64();
#endif
    // ===================================== phase 2: tuning =====================================
    // peak records: segment w of the slice holds cs.s_i[20 + w] records (the split pipeline uses segment 0 only)
    int np = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) np += cs.s_i[20 + w];
    float gmx = cs.s_f[0];
#pragma unroll
    for (int w = 1; w < kWarps; ++w) gmx = fmaxf(gmx, cs.s_f[w]);

    int tuning_idx = kTunings / 2;       // edges[50] == 0.0: librosa returns 0.0 for an empty pitch set
    float thr = 0.0f;
    int nsel = 0;
    if (kDebug) {                        // cs.s_i[16]: peaks whose fast-path shift or bin differs from the reference form
        if (tid == 0) cs.s_i[16] = 0;
        __syncthreads();
    }
    if (np > 0) {
        // ---- per-peak arithmetic of librosa.piptrack / pitch_tuning at full lane occupancy
        // Four peaks per thread and step, branch-free so that the four dependency chains interleave.  The two expensive
        // pieces have fast forms whose result is provably the reference one unless a guard fires:
        //   shift = float(-b/a): Newton quotient (relative error < 2^-50), accepted unless it lies within 2^-44 of a
        //           float32 rounding boundary or an operand leaves the normal float range; else redone at once with the
        //           IEEE division (peak_shift_exact; rare);
        //   bin:   residual from MUFU.LG2 of the mantissa (error < 1.5e-5 in 12*log2 against the reference form),
        //           accepted unless it falls within 0.25 % of a bin width of an edge of the 100-bin grid (which includes
        //           the +-0.5 wrap); else the peak is queued and redone after the loop (peak_bin_exact) by all threads.
        const bool in_smem = np <= kKeyCap;
        unsigned* keys = in_smem ? reinterpret_cast<unsigned*>(cs.s_ex) : sl.gKey;
        unsigned char* bins = in_smem ? reinterpret_cast<unsigned char*>(cs.s_ex + kKeyCap) : sl.gBin;
        // (peak index, pitch bits), kRedoCap entries, in the 1 KB of the warp tiles behind the shared-memory key and bin arrays
        uint2* redo_list = reinterpret_cast<uint2*>(cs.s_ex + kKeyCap + kKeyCap / 4);
        auto peaks = [&](auto SM) {
            constexpr bool kSmem = decltype(SM)::value;
            const float4 kDummy = make_float4(0.f, 1.f, 1.f, __int_as_float(64));      // harmless stand-in past the end
            // index -> record.  Dense form: one index over all segments, a thread's indices only grow, so it walks the
            // segment boundaries once.  kWarpSegs: index within the warp's own segment, keys land behind those of the
            // segments before it.
            const int first = kWarpSegs ? lane : tid, stride = kWarpSegs ? 32 : kThreads;
            const int count = kWarpSegs ? cs.s_i[20 + warp] : np;
            int off = 0;
            if constexpr (kWarpSegs) {
#pragma unroll
                for (int v = 0; v < kWarps; ++v) off += (v < warp) ? cs.s_i[20 + v] : 0;
            }
            const float4* seg = sl.gRec + static_cast<size_t>(warp) * sl.seg_cap;
            int seg_w = 0, seg_end = cs.s_i[20], seg_adj = 0;          // record of index i sits at gRec[i + seg_adj]
            auto fetch = [&](int i) -> float4 {
                if (i >= count) return kDummy;
                if constexpr (kWarpSegs) {
                    return seg[i];
                } else {
                    while (i >= seg_end) {
                        ++seg_w;
                        seg_adj += sl.seg_cap - cs.s_i[20 + seg_w - 1];
                        seg_end += cs.s_i[20 + seg_w];
                    }
                    return sl.gRec[i + seg_adj];
                }
            };
            (void)seg; (void)seg_w; (void)seg_end; (void)seg_adj;
            unsigned kor = 0u, kand = 0xffffffffu;
            float4 nxt[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) nxt[u] = fetch(first + u * stride);
            for (int i0 = first; i0 < count; i0 += 4 * stride) {
                float4 recs[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    recs[u] = nxt[u];
                    nxt[u] = fetch(i0 + (4 + u) * stride);                                // next step's records
                }
                float shift[4];
                bool redo = false;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float pm = recs[u].x, pc = recs[u].y, pp = recs[u].z;
                    const float sum = __fadd_rn(pp, pm);
                    const float dif = __fsub_rn(pp, pm);
                    const double a = static_cast<double>(sum) - 2.0 * static_cast<double>(pc);
                    const double b = static_cast<double>(dif) * 0.5;
                    const float af = static_cast<float>(a);
                    double r = static_cast<double>(rcp_approx(af));
                    r = fma(fma(-a, r, 1.0), r, r);
                    const double q1 = b * r;
                    const double q = -fma(fma(-q1, a, b), r, q1);
                    const unsigned qlo = static_cast<unsigned>(__double2loint(q)) & 0x1fffffffu;     // bits float32 drops
                    const unsigned qe = (static_cast<unsigned>(__double2hiint(q)) >> 20) & 0x7ffu;   // biased exponent
                    const unsigned ae = (__float_as_uint(af) >> 23) & 0xffu;
                    const bool zero = fabs(b) >= fabs(a);
                    const bool risky = ((qlo - 0x0fffff00u) < 0x200u) | (qe < 1023u - 100u) | ((ae - 27u) > 200u);
                    redo |= risky & !zero & (dif != 0.0f);
                    shift[u] = zero ? 0.0f : static_cast<float>(q);
                }
                if (redo) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) shift[u] = peak_shift_exact(recs[u].x, recs[u].y, recs[u].z);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float pm = recs[u].x, pc = recs[u].y, pp = recs[u].z;
                    const int k = __float_as_int(recs[u].w);
                    const float avg = __fsub_rn(pp, pm) * 0.5f;
                    const float dskew = __fmul_rn(__fmul_rn(0.5f, avg), shift[u]);
                    const unsigned key = fkey(__fadd_rn(pc, dskew));
                    const double pitch_d = (static_cast<double>(k) + static_cast<double>(shift[u])) *
                                           static_cast<double>(tb.sr) / static_cast<double>(kNfft);
                    const float pitch = static_cast<float>(pitch_d);
                    // pitch_tuning: mod(12*log2(f/27.5), 1) wrapped to [-0.5, 0.5), then its 0.01-wide bin.  Only the
                    // fractional part matters, so the exponent of f drops out: 12*log2(mantissa) - frac-preserving const
                    const unsigned pb = __float_as_uint(pitch);
                    const float mant = __uint_as_float((pb & 0x007fffffu) | 0x3f800000u);
                    const float w = fmaf(12.0f, lg2_approx(mant), -9.37631656229592f);    // 12*(log2(m) - (log2(27.5) - 4))
                    float res = w - floorf(w);
                    if (res >= 0.5f) res -= 1.0f;
                    const float uf = fmaf(res, 100.0f, 50.0f);
                    const float fl = floorf(uf);
                    const float fr = uf - fl;
                    int bin = max(0, min(kTunings - 1, static_cast<int>(fl)));
                    const int j = i0 + u * stride, i = off + j;       // index in the segment / dense index, key slot
                    // normal positive pitch and clear of the bin edges, else redo (NaN-safe: comparisons false -> redo)
                    const bool sure = (fr > 0.0025f) & (fr < 0.9975f) & ((pb - 0x00800000u) < 0x7f000000u);
                    if (!sure && j < count) {
                        const int slot = atomicAdd(&cs.s_i[17], 1);
                        if (slot < kRedoCap) redo_list[slot] = make_uint2(static_cast<unsigned>(i), pb);
                        else bin = peak_bin_exact(pitch, cs.s_edges);
                    }
                    if (j < count) {
                        kor |= key;
                        kand &= key;
                        if constexpr (kSmem) {
                            reinterpret_cast<unsigned*>(cs.s_ex)[i] = key;
                            reinterpret_cast<unsigned char*>(cs.s_ex + kKeyCap)[i] = static_cast<unsigned char>(bin);
                        } else {
                            sl.gKey[i] = key;
                            sl.gBin[i] = static_cast<unsigned char>(bin);
                        }
                    }
                }
            }
            kor = __reduce_or_sync(0xffffffffu, kor);
            kand = __reduce_and_sync(0xffffffffu, kand);
            if (lane == 0) {
                atomicOr(reinterpret_cast<unsigned*>(&cs.s_i[18]), kor);
                atomicAnd(reinterpret_cast<unsigned*>(&cs.s_i[19]), kand);
            }
        };
        if (in_smem) peaks(std::true_type{}); else peaks(std::false_type{});
        for (int i = tid; i < 256; i += kThreads) cs.s_hist[i] = 0;          // the select's first histogram
        __syncthreads();
        if constexpr (kWarpSegs && !kDebug && (SFX_DISCARD & 1)) {
            // the records are dead: drop their lines from L2 instead of letting them be written back to HBM (16 B per peak)
#pragma unroll 1
            for (int w = 0; w < kWarps; ++w)
                discard_l2_range(sl.gRec + static_cast<size_t>(w) * sl.seg_cap, static_cast<size_t>(cs.s_i[20 + w]) * sizeof(float4), tid);
        }
        FPROF_MARK(1);
        {
            const int nredo = min(cs.s_i[17], kRedoCap);
            for (int j = tid; j < nredo; j += kThreads) {
                const uint2 e = redo_list[j];
                bins[e.x] = static_cast<unsigned char>(peak_bin_exact(__uint_as_float(e.y), cs.s_edges));
            }
        }
        if (kDebug) {
            // every peak again with the reference forms only; keys / bins must be identical
            __syncthreads();
            int diff = 0;
            int seg_w = 0, seg_end = cs.s_i[20], seg_adj = 0;
            for (int i = tid; i < np; i += kThreads) {
                while (i >= seg_end) {
                    ++seg_w;
                    seg_adj += sl.seg_cap - cs.s_i[20 + seg_w - 1];
                    seg_end += cs.s_i[20 + seg_w];
                }
                const float4 rc = sl.gRec[i + seg_adj];
                const float sh = peak_shift_exact(rc.x, rc.y, rc.z);
                const float avg = __fsub_rn(rc.z, rc.x) * 0.5f;
                const unsigned key = fkey(__fadd_rn(rc.y, __fmul_rn(__fmul_rn(0.5f, avg), sh)));
                const double pitch_d = (static_cast<double>(__float_as_int(rc.w)) + static_cast<double>(sh)) *
                                       static_cast<double>(tb.sr) / static_cast<double>(kNfft);
                const int be = peak_bin_exact(static_cast<float>(pitch_d), cs.s_edges);
                diff += (key != keys[i]) || (be != static_cast<int>(bins[i]));
            }
            if (diff) atomicAdd(&cs.s_i[16], diff);
        }
        // (no barrier: the redo loop above writes bins, which nothing reads before the barriers of the select)
        FPROF_MARK(8);
        int cle = 0;
        const int r0 = (np - 1) >> 1;
        unsigned knext = 0u;
        bool has_next = false;
        // (histograms 1 and 2 of the select live in s_pool, free between the redo list above and the MFCC pooling)
        int* thist = nullptr;                            // a cleared histogram for the tuning bins, handed back by the select
        const unsigned ka = radix_select(keys, np, r0, cs.s_hist, reinterpret_cast<int*>(cs.s_pool), reinterpret_cast<int*>(cs.s_pool) + 256,
                                         cle, static_cast<unsigned>(cs.s_i[18]), static_cast<unsigned>(cs.s_i[19]), knext, has_next,
                                         thist);
        unsigned kb = ka;
        FPROF_MARK(9);
        if ((np & 1) == 0) {
            if (has_next) {
                kb = knext;                                  // the select ranked the key of rank r0 + 1 as well
            } else if (cle <= (np >> 1)) {
                // upper median = smallest key above ka
                if (tid == 0) cs.s_i[4] = static_cast<int>(0xffffffffu);
                __syncthreads();
                unsigned best = 0xffffffffu;
                for (int i0 = tid; i0 < np; i0 += 4 * kThreads) {
                    unsigned k4[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) k4[u] = (i0 + u * kThreads < np) ? keys[i0 + u * kThreads] : 0u;
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (k4[u] > ka && k4[u] < best) best = k4[u];
                }
                atomicMin(reinterpret_cast<unsigned*>(&cs.s_i[4]), best);
                __syncthreads();
                kb = static_cast<unsigned>(cs.s_i[4]);
                __syncthreads();
            }
        }
        FPROF_MARK(10);
        const float fa = fkey_inv(ka), fb = fkey_inv(kb);
        thr = ((np & 1) == 0) ? __fmul_rn(__fadd_rn(fa, fb), 0.5f) : fa;
        const unsigned kthr = fkey(thr);
        // histogram of the residual bins of peaks with mag >= median, into the cleared histogram the select handed back
        for (int i0 = tid; i0 < np; i0 += 4 * kThreads) {
            unsigned k4[4];
            int b4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool v = i0 + u * kThreads < np;
                k4[u] = v ? keys[i0 + u * kThreads] : 0u;
                b4[u] = v ? bins[i0 + u * kThreads] : 0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i0 + u * kThreads < np && k4[u] >= kthr) atomicAdd(&thist[b4[u]], 1);
        }
        fence_proxy_async_smem();       // last generic-proxy accesses of the tiles (keys, bins): the bank's async copy follows
        __syncthreads();
        {
            // first arg-max, by every warp for itself (no second barrier to publish one warp's answer)
            int bc = -1, bi = 1 << 20, tot = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int b = lane * 4 + q;
                if (b < kTunings) {
                    const int c = thist[b];
                    tot += c;
                    if (c > bc) { bc = c; bi = b; }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (oc > bc || (oc == bc && oi < bi)) { bc = oc; bi = oi; }
            }
            tuning_idx = bi;
            nsel = warp_sum_i(tot);
        }
    } else {
        fence_proxy_async_smem();       // phase 1's generic-proxy accesses of the tiles precede the bank's async copy
        __syncthreads();
    }
    if (kDebug) {
        if (p.dbg.clip_info && tid == 0) {
            float* ci = p.dbg.clip_info + static_cast<size_t>(clip) * 8;
            ci[0] = static_cast<float>(cs.s_edges[tuning_idx]);
            ci[1] = gmx; ci[2] = static_cast<float>(np); ci[3] = thr;
            ci[4] = static_cast<float>(nsel); ci[5] = static_cast<float>(T);
            ci[6] = static_cast<float>(cs.s_i[16]); ci[7] = 0.f;
        }
    }

    FPROF_MARK(2);
    // ===================================== phase 3a: MFCC ======================================
    // The tuning's FP16 hi/lo chroma bank (50 688 B) is staged into the now-free warp tiles by one TMA bulk copy
    // (cp.async.bulk, completes on an mbarrier) that runs underneath the MFCC pooling.
    __half* sW = reinterpret_cast<__half*>(cs.s_ex);       // [2][12][kP16Stride]  (fenced and behind a barrier: see above)
    if constexpr (!kUmmaTail) {
        if (tid == 0)
            bulk_copy_g2s(sW, tb.chroma16 + static_cast<size_t>(tuning_idx) * 2 * kChroma * kP16Stride,
                          2 * kChroma * kP16Stride * 2, cs.s_mbar);
    }
    FPROF_MARK(11);
    {
        const float clampv = __fsub_rn(gmx, 80.0f);
        // power_to_db's clamp max(L, gmax - 80) is the identity on every non-zero frame of most clips (it exists for the
        // zero tail load_audio pads short clips with, whose rows are the constant 10*log10(1e-10) exactly): phase 1 keeps
        // per-warp float64 sums of the unclamped rows of its non-zero frames and their minimum, and when that minimum is
        // not below the clamp the pooled mean follows from the sums and the count of all-zero frames without reading the
        // [T][128] rows back.  Sums of float32 values in float64 are exact here (53 bits hold 24-bit values over this
        // range), so both forms give the same bits; otherwise the rows are read (same arithmetic as ever).
        bool fast = false;
        if (cs.s_lm != nullptr) {
            float lmin = cs.s_lmin[tid & (kWarps * 32 - 1)];
            lmin = -warp_max(-lmin);
            if (lane == 0) cs.s_f[24 + warp] = lmin;
            __syncthreads();
            lmin = cs.s_f[24];
#pragma unroll
            for (int w = 1; w < kWarps; ++w) lmin = fminf(lmin, cs.s_f[24 + w]);
            fast = lmin >= clampv;
        }
        FPROF_MARK(12);
        if (fast) {
            if (tid < 128) {
                float nz = cs.s_f[16];
#pragma unroll
                for (int w = 1; w < kWarps; ++w) nz += cs.s_f[16 + w];
                const float lm0 = 3.01029995663981195f * __log2f(1e-10f);
                double a = static_cast<double>(nz) * static_cast<double>(fmaxf(lm0, clampv));
#pragma unroll
                for (int w = 0; w < kWarps; ++w) a += cs.s_lm[w * kMels + tid];
                cs.s_pool[tid] = a / static_cast<double>(T);
            }
        } else {
            if (tid < 256) {
                const int m = tid & 127, h = tid >> 7;
                double a = 0.0;
                for (int t = h; t < T; t += 2) a += static_cast<double>(fmaxf(sl.gL[static_cast<size_t>(t) * kMels + m], clampv));
                cs.s_pool[h * 128 + m] = a;
            }
            __syncthreads();
            if (tid < 128) cs.s_pool[tid] = (cs.s_pool[tid] + cs.s_pool[128 + tid]) / static_cast<double>(T);
        }
        __syncthreads();
        if constexpr (kWarpSegs && (SFX_DISCARD & 2)) discard_l2_range(sl.gL, static_cast<size_t>(T) * kMels * sizeof(float), tid);   // rows pooled: dead
        FPROF_MARK(13);
        {
            // DCT-II of the pooled log-mel vector: coefficient k by a pair of threads (bands 0..63 and 64..127, the two
            // partial sums added in that order), which halves the dependent float64 FMA chain on the tail's critical path
            // (six threads per coefficient through shared memory measured 0.4 % slower: one more barrier)
            const int k = tid >> 1, h = tid & 1;
            double d = 0.0;
            if (k < p.n_mfcc)
                for (int q = 64 * h; q < 64 * h + 64; ++q) d = fma(tb.dctT[q * kMels + k], cs.s_pool[q], d);
            const double o = __shfl_xor_sync(0xffffffffu, d, 1);
            if (k < p.n_mfcc && h == 0) out[k] = static_cast<float>(d + o);
        }
    }
    FPROF_MARK(3);
    if constexpr (!kUmmaTail) {
        mbar_wait(cs.s_mbar, bank_parity);                 // chroma bank has landed in shared memory
        bank_parity ^= 1u;
    }
    FPROF_MARK(4);

    // ===================================== phase 3b: chroma ====================================
    // raw[c][t] = sum_k W[c][k] |X|^2[k][t] on the tensor cores: m16n8k16 FP16 MMAs with FP32 accumulators.  A = bank
    // (hi and 2^11*lo halves, separate accumulators), B = the frame's scaled FP16 |X|^2 row.  One unit = 8 frames x
    // 512 bins = 16 steps of 1 LDG.128 + 4 LDS.128 + 4 MMA; lane (g = lane/4, t4 = lane%4) feeds frame g's bins
    // k0+8*t4..+7 and rows g, g+8 of the bank.  Units are dealt round-robin to the warps, the two K-halves of a tile
    // are added in a fixed order, then each frame is normalised by its max (librosa norm=inf).  The clip's last,
    // incomplete tile is split by steps instead (below).
    {
        const int g = lane >> 2, t4 = lane & 3;
        (void)g; (void)t4;
        double csum[kChroma];                                        // per-thread sums over its frames
#pragma unroll
        for (int c = 0; c < kChroma; ++c) csum[c] = 0.0;
        float wny[kChroma];
#pragma unroll
        for (int c = 0; c < kChroma; ++c) wny[c] = __ldg(tb.chroma_ny + tuning_idx * kChroma + c);
      if constexpr (kUmmaTail) {
        // tcgen05 form: D[128 frames x 32] (+)= A[128 frames x 64 bins] . B[32 x 64 bins]^T per 64-bin K block, FP16 operands,
        // FP32 accumulator in tensor memory; columns 0..11 = W_hi . |X|^2, 16..27 = 2^11 W_lo . |X|^2.  Phase 1 stored the
        // rows as the shared-memory images of the A blocks, the host built the B images of every tuning, so an operand block
        // is ONE contiguous bulk copy (cp.async.bulk, no tensor map) into a 3-stage ring in the warp tiles; thread 0 issues
        // copies and MMAs (4 x K=16 per block), tcgen05.commit frees a stage / publishes the accumulator, warps 0-3 read
        // their 32 tensor-memory lanes (one frame per thread) and normalise.
        float lastnz = cs.s_f[8];
#pragma unroll
        for (int w = 1; w < kWarps; ++w) lastnz = fmaxf(lastnz, cs.s_f[8 + w]);
        const int Tc = min(T, static_cast<int>(lastnz) + 1);           // frames behind the last non-zero one add exactly 0
        // stage ring: first 1 KB boundary inside the warp tiles, 3 x (16 KB A | 4 KB B)
        const unsigned ring = (smem_u32(cs.s_ex) + 1023u) & ~1023u;
        unsigned char* ring_g = reinterpret_cast<unsigned char*>(cs.s_ex) + (ring - smem_u32(cs.s_ex));
        constexpr unsigned kStageBytes = 20480u, kIdesc = (1u << 4) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
        const unsigned char* bimg = tb.chroma_umma + static_cast<size_t>(tuning_idx) * (16 * 4096);
        const unsigned char* aimg = reinterpret_cast<const unsigned char*>(sl.gP16);
        const int natoms = (p.Tmax + 7) >> 3;
        // full and large partial tiles on the tensor cores; a remainder of fewer than 32 frames (3 s clips: frames 128, 129)
        // is not worth a 16-block operand stream and goes through the FP32 pipes below
        const int n_umma = (Tc >> 7) + ((Tc & 127) >= 32 ? 1 : 0);
        for (int tile = 0; tile < n_umma; ++tile) {
            const int na_alloc = min(16, natoms - 16 * tile);                        // atoms the tile owns in the scratch image
            const int na = min(na_alloc, (Tc - 128 * tile + 7) >> 3);                // atoms that hold frames of this clip
            if (tid == 0) {
                const unsigned c0 = us->blocks;
                auto issue_copy = [&](int i) {
                    const unsigned c = c0 + i, st = c % 3u;
                    if (c >= 3u) mbar_wait(cs.s_ubar + 3 + st, ((c / 3u) - 1u) & 1u);   // the MMAs that read this stage are done
                    mbar_expect_tx(cs.s_ubar + st, static_cast<unsigned>(na) * 1024u + 4096u);
                    bulk_copy_g2s_nx(ring_g + st * kStageBytes, aimg + static_cast<size_t>(tile) * (16 * 16 * 1024) +
                                     static_cast<size_t>(i) * na_alloc * 1024, static_cast<unsigned>(na) * 1024u, cs.s_ubar + st);
                    bulk_copy_g2s_nx(ring_g + st * kStageBytes + 16384, bimg + static_cast<size_t>(i) * 4096, 4096u, cs.s_ubar + st);
                };
                issue_copy(0);
                issue_copy(1);
                issue_copy(2);
                for (int i = 0; i < 16; ++i) {
                    const unsigned c = c0 + i, st = c % 3u;
                    mbar_wait(cs.s_ubar + st, (c / 3u) & 1u);
                    tc_fence_after();
                    const unsigned long long ad = umma_desc_sw128(ring + st * kStageBytes);
                    const unsigned long long bd = umma_desc_sw128(ring + st * kStageBytes + 16384u);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16(cs.tmem, ad + 2ull * k, bd + 2ull * k, kIdesc, (i | k) != 0 ? 1u : 0u);
                    umma_commit(cs.s_ubar + 3 + st);
                    if (i + 3 < 16) issue_copy(i + 3);            // into the stage these MMAs are reading: waits for their commit
                }
                umma_commit(cs.s_ubar + 6);
                mbar_wait(cs.s_ubar + 6, us->tiles & 1u);              // the tile's accumulator is complete (one poller:
            }                                                          // everybody else sleeps at the barrier below)
            us->blocks += 16u;
            us->tiles += 1u;
            __syncthreads();
            tc_fence_after();
            if (warp < 4) {
                float v[32];
                tmem_ld32(cs.tmem + (static_cast<unsigned>(warp * 32) << 16), v);
                const int f = tile * 128 + warp * 32 + lane;
                if (f < Tc) {
                    const float pn = sl.gNy[f];                            // scaled Nyquist bin
                    float raw[kChroma];
                    float mx = 0.0f;
#pragma unroll
                    for (int c = 0; c < kChroma; ++c) {
                        raw[c] = fmaf(wny[c], pn, fmaf(v[16 + c], 1.0f / 2048.0f, v[c]));
                        mx = fmaxf(mx, fabsf(raw[c]));
                    }
                    // librosa.util.normalize: lengths below tiny(float32) are replaced by 1 (in unscaled units)
                    const float inv_s = sl.gInvS[f];
                    const bool small = mx * inv_s < FLT_MIN;
#pragma unroll
                    for (int c = 0; c < kChroma; ++c)
                        csum[c] += static_cast<double>(small ? raw[c] * inv_s : __fdiv_rn(raw[c], mx));
                }
            }
            tc_fence_before();
            __syncthreads();                                           // every reader is done before the next tile overwrites
        }
        // remainder frames on the FP32 pipes: work item = (frame, chroma class, 128-bin slice), one per thread (3 s clips:
        // 2 frames x 12 x 8 = 192 items), loads batched four 8-bin chunks deep, 8-lane shuffle tree over the slices, the
        // (frame, class) sums meet in shared memory and one thread per frame normalises
        {
            const int f0 = n_umma * 128, nrem = Tc - f0;               // 0 <= nrem < 32
            const __half* bank = reinterpret_cast<const __half*>(tb.chroma16) + static_cast<size_t>(tuning_idx) * 2 * kChroma * kP16Stride;
            float* s_rem = reinterpret_cast<float*>(cs.s_pool);        // [nrem][12] (the pooled log-mel vector is consumed)
            const int nitems = nrem * kChroma * 8;
            for (int base = 0; base < nitems; base += kThreads) {      // CTA-uniform trip count (shuffles inside)
                const int item = base + tid;
                const bool valid = item < nitems;
                const int part = item & 7, c = (item >> 3) % kChroma, fi = valid ? item / (8 * kChroma) : 0;
                const int f = f0 + fi;
                float acc = 0.0f;
                if (valid) {
                    const __half* wh = bank + c * kP16Stride + 128 * part;
                    const __half* wl = bank + (kChroma + c) * kP16Stride + 128 * part;
#pragma unroll 1
                    for (int ch0 = 0; ch0 < 16; ch0 += 4) {
                        uint4 pv[4], hv[4], lv[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int bin0 = 128 * part + 8 * (ch0 + q);
                            pv[q] = *reinterpret_cast<const uint4*>(aimg + umma_p16_offset(p.Tmax, f, bin0 >> 6, (bin0 >> 3) & 7));
                            hv[q] = *reinterpret_cast<const uint4*>(wh + 8 * (ch0 + q));
                            lv[q] = *reinterpret_cast<const uint4*>(wl + 8 * (ch0 + q));
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const __half2* ph = reinterpret_cast<const __half2*>(&pv[q]);
                            const __half2* hh = reinterpret_cast<const __half2*>(&hv[q]);
                            const __half2* hl = reinterpret_cast<const __half2*>(&lv[q]);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 x = __half22float2(ph[e]), a = __half22float2(hh[e]), b2 = __half22float2(hl[e]);
                                acc = fmaf(fmaf(b2.x, 1.0f / 2048.0f, a.x), x.x, acc);
                                acc = fmaf(fmaf(b2.y, 1.0f / 2048.0f, a.y), x.y, acc);
                            }
                        }
                    }
                }
                acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                acc += __shfl_xor_sync(0xffffffffu, acc, 4);
                if (valid && part == 0) s_rem[fi * kChroma + c] = acc;
            }
            __syncthreads();
            if (tid < nrem) {
                const int f = f0 + tid;
                const float pn = sl.gNy[f];
                float raw[kChroma];
                float mx = 0.0f;
#pragma unroll
                for (int c = 0; c < kChroma; ++c) {
                    raw[c] = fmaf(wny[c], pn, s_rem[tid * kChroma + c]);
                    mx = fmaxf(mx, fabsf(raw[c]));
                }
                const float inv_s = sl.gInvS[f];
                const bool small = mx * inv_s < FLT_MIN;
#pragma unroll
                for (int c = 0; c < kChroma; ++c)
                    csum[c] += static_cast<double>(small ? raw[c] * inv_s : __fdiv_rn(raw[c], mx));
            }
            __syncthreads();
        }
      } else {
        float* part2 = cs.s_ex + (2 * kChroma * kP16Stride) / 2;        // [kChromaTiles][2][96] floats after the bank
        float* part3 = part2 + kChromaTiles * 2 * 96;                   // [kWarps][96]: the incomplete tile, split by steps
        // Frames behind the clip's last non-zero frame (the zero tail load_audio pads short clips with) have all-zero
        // |X|^2 rows and add exactly 0 to every chroma sum: the projection stops at Tc = that frame + 1 (cs.s_f[8 + w] =
        // last non-zero frame seen by warp w, -1 if none); the mean below still divides by T.
        float lastnz = cs.s_f[8];
#pragma unroll
        for (int w = 1; w < kWarps; ++w) lastnz = fmaxf(lastnz, cs.s_f[8 + w]);
        const int Tc = min(T, static_cast<int>(lastnz) + 1);
        // full 8-frame tiles, 16 per pass (= 16 units, 2 per warp); the clip's last, incomplete tile (Tc % 8 frames) rides
        // along with the last pass: its 32 steps are split evenly over the 8 warps (4 each), so that 130 frames cost every
        // warp 2 units + 4 steps, one barrier and one normalisation instead of a pass of their own
        const int nfull = Tc >> 3, rem = Tc & 7;
        const int npass = max(1, (nfull + kChromaTiles - 1) / kChromaTiles);
        // normalisation: a pair of threads per frame, six classes each (half = tid & 1), the frame maximum by one shuffle
        const int half = tid & 1;
        double csum6[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};              // per-thread sums over its frames, classes 6*half..+5
        for (int pass = 0; pass < npass; ++pass) {
            const int tile0 = pass * kChromaTiles;
            const int nt = min(kChromaTiles, nfull - tile0);           // 0 when the clip has no full tile at all
            const bool tail_pass = (pass == npass - 1) && rem;
            // unit = pair of tiles x K-half: every bank fragment read from shared memory feeds the MMAs of two tiles (16
            // frames), which halves the bank's shared-memory traffic (it was 64 wavefronts per frame, a tenth of the kernel's)
            const int npair = (nt + 1) >> 1;
            for (int u = warp; u < 2 * npair; u += kWarps) {
                const int pr = u >> 1, kh = u & 1;
                const bool vB = 2 * pr + 1 < nt;                      // an odd tile count leaves the last pair half empty
                const int fA = (tile0 + 2 * pr) * 8 + g, fB = fA + (vB ? 8 : 0);
                const __half* prowA = sl.gP16 + static_cast<size_t>(fA) * kP16Row + kh * 512 + 8 * t4;
                const __half* prowB = sl.gP16 + static_cast<size_t>(fB) * kP16Row + kh * 512 + 8 * t4;
                const int r1 = (g < 4) ? g + 8 : g;                  // bank rows 12..15 do not exist
                const __half* whi0 = sW + g * kP16Stride + kh * 512 + 8 * t4;
                const __half* whi1 = sW + r1 * kP16Stride + kh * 512 + 8 * t4;
                const __half* wlo0 = whi0 + kChroma * kP16Stride;
                const __half* wlo1 = whi1 + kChroma * kP16Stride;
                // one accumulator set per tile and bank half: four MMA chains in flight; lanes
                // g >= 4 feed bank row g again as the non-existent rows 12..15, whose D rows are never read
                float acc[2][4] = {}, acl[2][4] = {};
#pragma unroll 1
                for (int kb0 = 0; kb0 < 16; kb0 += 4) {
                    uint4 pa[4], pb[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        pa[i] = *reinterpret_cast<const uint4*>(prowA + (kb0 + i) * 32);
                        pb[i] = *reinterpret_cast<const uint4*>(prowB + (kb0 + i) * 32);
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int o = (kb0 + i) * 32;
                        const uint4 h0 = *reinterpret_cast<const uint4*>(whi0 + o);
                        const uint4 h1 = *reinterpret_cast<const uint4*>(whi1 + o);
                        const uint4 l0 = *reinterpret_cast<const uint4*>(wlo0 + o);
                        const uint4 l1 = *reinterpret_cast<const uint4*>(wlo1 + o);
                        mma_f16(acc[0], h0.x, h1.x, h0.y, h1.y, pa[i].x, pa[i].y);
                        mma_f16(acc[1], h0.x, h1.x, h0.y, h1.y, pb[i].x, pb[i].y);
                        mma_f16(acl[0], l0.x, l1.x, l0.y, l1.y, pa[i].x, pa[i].y);
                        mma_f16(acl[1], l0.x, l1.x, l0.y, l1.y, pb[i].x, pb[i].y);
                        mma_f16(acc[0], h0.z, h1.z, h0.w, h1.w, pa[i].z, pa[i].w);
                        mma_f16(acc[1], h0.z, h1.z, h0.w, h1.w, pb[i].z, pb[i].w);
                        mma_f16(acl[0], l0.z, l1.z, l0.w, l1.w, pa[i].z, pa[i].w);
                        mma_f16(acl[1], l0.z, l1.z, l0.w, l1.w, pb[i].z, pb[i].w);
                    }
                }
                // D fragment: [0..1] = (chroma g, frames 2*t4, 2*t4+1), [2..3] = (chroma g+8, same frames)
                constexpr float kLo = 1.0f / 2048.0f;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (h == 1 && !vB) break;
                    float* dst = part2 + ((2 * pr + h) * 2 + kh) * 96;
                    *reinterpret_cast<float2*>(dst + g * 8 + 2 * t4) =
                        make_float2(fmaf(acl[h][0], kLo, acc[h][0]), fmaf(acl[h][1], kLo, acc[h][1]));
                    if (g < 4)
                        *reinterpret_cast<float2*>(dst + (g + 8) * 8 + 2 * t4) =
                            make_float2(fmaf(acl[h][2], kLo, acc[h][2]), fmaf(acl[h][3], kLo, acc[h][3]));
                }
            }
            if (tail_pass) {
                const int f = nfull * 8 + g;
                const bool valid = f < Tc;
                const int k0 = warp * 128 + 8 * t4;                      // steps 4*warp .. 4*warp+3
                const __half* prow = sl.gP16 + static_cast<size_t>(valid ? f : 0) * kP16Row + k0;
                const int r1 = (g < 4) ? g + 8 : g;
                const __half* whi0 = sW + g * kP16Stride + k0;
                const __half* whi1 = sW + r1 * kP16Stride + k0;
                const __half* wlo0 = whi0 + kChroma * kP16Stride;
                const __half* wlo1 = whi1 + kChroma * kP16Stride;
                float acc[4] = {0.f, 0.f, 0.f, 0.f}, acl[4] = {0.f, 0.f, 0.f, 0.f};
                uint4 pv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) pv[i] = valid ? *reinterpret_cast<const uint4*>(prow + i * 32) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int o = i * 32;
                    const uint4 h0 = *reinterpret_cast<const uint4*>(whi0 + o);
                    const uint4 h1 = *reinterpret_cast<const uint4*>(whi1 + o);
                    const uint4 l0 = *reinterpret_cast<const uint4*>(wlo0 + o);
                    const uint4 l1 = *reinterpret_cast<const uint4*>(wlo1 + o);
                    mma_f16(acc, h0.x, h1.x, h0.y, h1.y, pv[i].x, pv[i].y);
                    mma_f16(acc, h0.z, h1.z, h0.w, h1.w, pv[i].z, pv[i].w);
                    mma_f16(acl, l0.x, l1.x, l0.y, l1.y, pv[i].x, pv[i].y);
                    mma_f16(acl, l0.z, l1.z, l0.w, l1.w, pv[i].z, pv[i].w);
                }
                constexpr float kLo = 1.0f / 2048.0f;
                float* dst = part3 + warp * 96;
                *reinterpret_cast<float2*>(dst + g * 8 + 2 * t4) = make_float2(fmaf(acl[0], kLo, acc[0]), fmaf(acl[1], kLo, acc[1]));
                if (g < 4)
                    *reinterpret_cast<float2*>(dst + (g + 8) * 8 + 2 * t4) = make_float2(fmaf(acl[2], kLo, acc[2]), fmaf(acl[3], kLo, acc[3]));
            }
            __syncthreads();
            const int nfr = nt * 8 + (tail_pass ? rem : 0);             // frames to normalise in this pass (<= 135)
            for (int fl0 = 0; fl0 < nfr; fl0 += kThreads / 2) {         // (CTA-uniform trip count: the shuffle below is full-warp)
                const int fl = fl0 + (tid >> 1);
                const bool live = fl < nfr;
                const int f = tile0 * 8 + (live ? fl : 0);
                const float pn = sl.gNy[f];                              // scaled Nyquist bin
                const float inv_s = sl.gInvS[f];
                float raw[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                float mx = 0.0f;
                if (!live) {
                } else if (fl < nt * 8) {
                    const float* q = part2 + (fl >> 3) * 192 + (fl & 7) + 48 * half;
#pragma unroll
                    for (int c = 0; c < 6; ++c) {
                        raw[c] = fmaf(wny[6 * half + c], pn, q[c * 8] + q[96 + c * 8]);
                        mx = fmaxf(mx, fabsf(raw[c]));
                    }
                } else {
                    const float* q = part3 + (fl - nt * 8) + 48 * half;   // the 8 step-quarters are added in warp order
#pragma unroll
                    for (int c = 0; c < 6; ++c) {
                        float v = q[c * 8];
#pragma unroll
                        for (int w = 1; w < kWarps; ++w) v += q[w * 96 + c * 8];
                        raw[c] = fmaf(wny[6 * half + c], pn, v);
                        mx = fmaxf(mx, fabsf(raw[c]));
                    }
                }
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                // librosa.util.normalize: lengths below tiny(float32) are replaced by 1 (in unscaled units)
                const bool small = mx * inv_s < FLT_MIN;
                if (live) {
#pragma unroll
                    for (int c = 0; c < 6; ++c)
                        csum6[c] += static_cast<double>(small ? raw[c] * inv_s : __fdiv_rn(raw[c], mx));
                }
            }
            if (pass + 1 < npass) __syncthreads();          // every reader is done before the next pass overwrites the partials
        }
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            double v = csum6[c];
#pragma unroll
            for (int o = 16; o > 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane < 2) cs.s_wacc[warp * 16 + 3 + 6 * lane + c] = v;
        }
      }
      if constexpr (kUmmaTail) {
#pragma unroll
        for (int c = 0; c < kChroma; ++c) {
            double v = csum[c];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) cs.s_wacc[warp * 16 + 3 + c] = v;
        }
      }
    }
    __syncthreads();
    if constexpr (kWarpSegs && !kUmmaTail && (SFX_DISCARD & 4)) discard_l2_range(sl.gP16, static_cast<size_t>(T) * kP16Row * sizeof(__half), tid);   // projected: dead

    FPROF_MARK(5);
    // ===================================== epilogue: pooled row ================================
    if (warp == 1) {
        // pooled rms from the hop energies (frame t spans hops t-2 .. t+1; hops outside [0, T) are zero padding)
        double a = 0.0;
        for (int t = lane; t < T; t += 32) {
            float e = (t >= 2) ? sl.gE[t - 2] : 0.0f;
            e += (t >= 1) ? sl.gE[t - 1] : 0.0f;
            e += sl.gE[t];
            e += (t + 1 < T) ? sl.gE[t + 1] : 0.0f;
            const float r = sqrtf(e * (1.0f / kNfft));
            a += static_cast<double>(r);
            if (kDebug) {
                if (p.dbg.frame_feat && t < p.dbg.T_dbg)
                    p.dbg.frame_feat[(static_cast<size_t>(clip) * p.dbg.T_dbg + t) * 4 + 2] = r;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) out[p.n_mfcc + 15] = static_cast<float>(a / static_cast<double>(T));
    }
    if (tid < 15) {
        const double invT = 1.0 / static_cast<double>(T);
        if (tid < kChroma) {
            double v = 0.0;
            for (int w = 0; w < kWarps; ++w) v += cs.s_wacc[w * 16 + 3 + tid];
            out[p.n_mfcc + tid] = static_cast<float>(v * invT);
        } else if (tid == 12) {
            long long z = 0;
            for (int w = 0; w < kWarps; ++w) z += cs.s_i[8 + w];
            out[p.n_mfcc + 12] = static_cast<float>(static_cast<double>(z) / (static_cast<double>(kNfft) * T));
        } else {
            double v = 0.0;
            for (int w = 0; w < kWarps; ++w) v += cs.s_wacc[w * 16 + (tid - 13)];
            out[p.n_mfcc + tid] = static_cast<float>(v * invT);
        }
    }
}

}  // namespace sfx
