// sfx_stream.cu -- the third pipeline of the extractor: one persistent 16-warp CTA per SM in which the per-clip tail never
// idles the FFT warps.
//
// The fused kernel (sfx_kernels.cu) runs a clip's tail (tuning estimate, MFCC, chroma, pooled row) with all 8 warps of its
// CTA: the tail holds ~12 % of the instructions but 25-40 % of a CTA's time, because it is a chain of scratch round trips
// and CTA barriers, and it only overlaps with FFT work of the *other* co-resident CTA.  Here every warp is an independent
// worker that pulls items from a CTA-local scheduler (a few words of shared memory, lock-free on the hot path: one 64-bit
// atomicAdd hands out a frame):
//   FRAME(slot, t)  phase 1 of STFT frame t of the clip in `slot` (process_frame<., kModeStream>: the code of the other
//                   pipelines); frames of the next clip are handed out as soon as the current clip's last frame has been
//                   *taken*, so there is no per-clip barrier and no round imbalance (130 frames over 8 warps);
//   TAIL(slot)      phases 2-3 of a clip whose frames are all done, executed by ONE warp from start to end
//                   (clip_tail_warp below: no CTA barrier anywhere).  A tail is latency-bound whoever runs it; run by one
//                   warp it costs one warp's time instead of eight, and the other 15 keep transforming frames.
// A CTA owns kSlots scratch slices; a slot is free -> frames being handed out / run -> ready (tail pending) -> tail -> free.  Per-frame descriptors are stored in
// the slice and summed by the tail warp in frame order, every frame owns a fixed segment of the slot's peak-record buffer,
// so the pooled row does not depend on which warp ran which frame.
//
// Chroma in the tail: raw = W . |X|^2 on the tensor cores (m16n8k16, FP16 hi/lo bank, FP32 accumulate) with the bank read
// straight from L2 (all 100 banks are 5 MB and resident): a bank fragment is applied to 32 frames (4 MMA N-tiles) per load,
// so no shared-memory copy of the bank is needed and any number of tails can run side by side.
#include "sfx_phases.cuh"

#ifndef SFX_TAIL_SKIP          // experiment: bit 0 skip phase 2, bit 1 skip MFCC, bit 2 skip chroma (rows are then garbage)
#define SFX_TAIL_SKIP 0
#endif
#ifndef SFX_TAIL_CG            // experiment: 1 = the tail's streaming loads bypass L1 (ld.global.cg)
#define SFX_TAIL_CG 0
#endif
#ifndef SFX_STREAM_SLOTS
#define SFX_STREAM_SLOTS 6
#endif

namespace sfx {

constexpr int kSW = kStreamWarps;
constexpr int kSThreads = kSW * 32;
constexpr int kSlots = SFX_STREAM_SLOTS;

// tail warp's use of its own 2112-float tile
constexpr int kTHist = 0;            // int[256]  radix-select / tuning histogram; later the pooled log-mel means (128 doubles)
constexpr int kTRedo = 256;          // uint2[64] peaks whose residual bin is redone with the reference form
constexpr int kTRedoCap = 64;
constexpr int kTKeys = 384;          // u32[kTKeyCap] keys, then u8[kTKeyCap] bins
constexpr int kTKeyCap = ((kExFloats - kTKeys) * 4 / 5) & ~3;

enum : int { kWorkExit = 0, kWorkWait = 1, kWorkFrame = 2, kWorkTail = 3, kWorkBad = 4 };

struct Sched {                       // shared memory, every field accessed through volatile or atomics
    unsigned hot;                    // frame ticket [gen:8 | t:24] of the clip being handed out: atomicAdd(&hot, 1) returns a
                                     // consistent (gen, t); ring_T / ring_slot[gen & 7] describe generation gen (valid until
                                     // 8 more clips have been opened); t < T means "frame t of that clip is yours"
    int ring_slot[8], ring_T[8];
    int ready_mask;                  // bit s: all frames of slot s are done, its tail is up for grabs
    int free_mask;                   // bit s: slot s is free
    int active;                      // slots that are not free
    int qdone;                       // the batch's clip queue is exhausted
    int opener;                      // spin lock of the (rare) "put the next clip into a free slot" path
    int clip[kSlots], T[kSlots], done[kSlots];
    int npk[kSlots];                 // peak records appended to the slot's dense record array so far
    long long n[kSlots];
};
static_assert(kSlots >= 2 && kSlots <= 8, "slot masks are 8 bits wide");

#ifdef SFX_STREAM_DIAG
// cycle counters per (CTA, warp): 0 frames, 1 tails, 2 waiting, 3 scheduler calls, 4 frames run, 5 tails run,
// 6 tail: sums + phase 2 per-peak loop, 7 tail: median + histogram, 8 tail: MFCC, 9 tail: chroma, 10 tail: epilogue, 11 peaks
__device__ long long g_prof[148 * 16 * 16];
#define PROF_ADD(k, v) do { if (lane == 0) g_prof[(blockIdx.x * 16 + (threadIdx.x >> 5)) * 16 + (k)] += (v); } while (0)
extern "C" int sfx_stream_prof(long long* out, int n, int reset) {
    int rc = static_cast<int>(cudaMemcpyFromSymbol(out, g_prof, sizeof(long long) * n));
    if (reset) {
        static long long zeros[148 * 16 * 16];
        rc |= static_cast<int>(cudaMemcpyToSymbol(g_prof, zeros, sizeof(zeros)));
    }
    return rc;
}
__device__ int g_diag[148 * 16 * 24];
__device__ void diag_dump(Sched* sc, int warp, int reason) {
    volatile Sched* v = sc;
    int* d = g_diag + (blockIdx.x * 16 + warp) * 24;
    d[0] = reason; d[1] = v->opener; d[2] = v->qdone; d[3] = v->ready_mask; d[4] = v->free_mask; d[5] = v->active;
    d[6] = static_cast<int>(v->hot >> 24); d[7] = static_cast<int>(v->hot & 0xffffffu);
    for (int s = 0; s < kSlots && s < 4; ++s) { d[11 + s] = v->done[s]; d[15 + s] = v->T[s]; d[19 + s] = v->clip[s]; }
}
extern "C" int sfx_stream_diag(int* out, int n) {
    return static_cast<int>(cudaMemcpyFromSymbol(out, g_diag, sizeof(int) * n));
}
#endif

#ifndef SFX_STREAM_DIAG
#define PROF_ADD(k, v) do { } while (0)
#endif

struct StreamSlice {
    __half* gP16; float* gL; float* gFv; float4* gRec; unsigned* gKey; unsigned char* gBin;
};

// layout of a slot: FP16 |X|^2 rows | log-mel rows | per-frame value records | peak records (max_pk per frame) | keys | bins
__device__ __forceinline__ StreamSlice stream_slice(unsigned char* base, int Tmax, int max_pk) {
    StreamSlice s;
    s.gP16 = reinterpret_cast<__half*>(base);
    s.gL = reinterpret_cast<float*>(s.gP16 + static_cast<size_t>(Tmax) * kP16Row);
    s.gFv = s.gL + static_cast<size_t>(Tmax) * kMels;
    s.gRec = reinterpret_cast<float4*>(s.gFv + static_cast<size_t>(Tmax) * kFvStride);
    s.gKey = reinterpret_cast<unsigned*>(s.gRec + static_cast<size_t>(Tmax) * max_pk);
    s.gBin = reinterpret_cast<unsigned char*>(s.gKey + static_cast<size_t>(Tmax) * max_pk);
    return s;
}

// ------------------------------------------------------------------------------------------------ one-warp histogram update
// (measured: aggregating equal bins with match.any and a plain read-modify-write by the lowest lane is slower than the
// shared-memory atomic -- noise clips 1.14 -> 0.94 M clips/s)
static __device__ __forceinline__ void hist_add(int* hist, const bool valid, const unsigned bin, const int) {
    if (valid) atomicAdd(&hist[bin], 1);
}

// ------------------------------------------------------------------------------------------------ one-warp radix select
// key of ascending rank r among keys[0..np); count_le = elements <= that key.  Same scheme as radix_select (only the bits
// in which the keys differ, 8 per pass), executed by one warp; hist = int[256] in the warp's tile.
static __device__ __forceinline__ unsigned radix_select_warp(const unsigned* keys, int np, int r, int* hist, int& count_le,
                                                             unsigned kor, unsigned kand, int lane) {
    const unsigned diff = kor ^ kand;
    int remaining = 32 - __clz(diff);
    unsigned mask = remaining >= 32 ? 0u : ~((1u << remaining) - 1u);
    unsigned prefix = kand & mask;
    int less = 0, equal = np;
    while (remaining > 0) {
        const int width = min(8, remaining);
        const int shift = remaining - width;
        const unsigned bmask = (1u << width) - 1u;
#pragma unroll
        for (int q = 0; q < 8; ++q) hist[lane + 32 * q] = 0;
        __syncwarp();
        for (int ib = 0; ib < np; ib += 128) {                 // warp-uniform trip count (match / syncwarp inside)
            const int i0 = ib + lane;
            unsigned k[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) k[u] = (i0 + 32 * u < np) ? keys[i0 + 32 * u] : 0u;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool valid = i0 + 32 * u < np && (k[u] & mask) == prefix;
                hist_add(hist, valid, (k[u] >> shift) & bmask, lane);
            }
        }
        __syncwarp();
        int loc[8], sum = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) { loc[q] = hist[lane * 8 + q]; sum += loc[q]; }
        int inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        const int exc = inc - sum;
        const unsigned bal = __ballot_sync(0xffffffffu, inc > r);
        const int L = __ffs(bal) - 1;
        int sel = 0, cum = 0, eq = 0;
        {
            const int rr = r - exc;
            bool found = false;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (!found) {
                    if (cum + loc[u] > rr) { found = true; sel = u; eq = loc[u]; }
                    else cum += loc[u];
                }
            }
        }
        const int bucket = __shfl_sync(0xffffffffu, lane * 8 + sel, L);
        const int below = __shfl_sync(0xffffffffu, exc + cum, L);
        equal = __shfl_sync(0xffffffffu, eq, L);
        r -= below;
        less += below;
        prefix |= static_cast<unsigned>(bucket) << shift;
        mask |= bmask << shift;
        remaining = shift;
        __syncwarp();
    }
    count_le = less + equal;
    return prefix;
}

template <class T>
static __device__ __forceinline__ T tail_ld(const T* ptr) {
#if SFX_TAIL_CG
    return __ldcg(ptr);
#else
    return *ptr;
#endif
}
static __device__ __forceinline__ uint4 tail_ldc(const uint4* ptr) {     // constant table
#if SFX_TAIL_CG
    return __ldcg(ptr);
#else
    return __ldg(ptr);
#endif
}

// One frame's contribution to the pooled chroma of rows g and g+8: librosa.util.normalize(norm=inf) -- raw / max|raw| over
// the 12 classes, with lengths below tiny(float32) (in unscaled units) replaced by 1.  Not inlined: the IEEE division is
// ~40 instructions with a slow path, and the tail's code size is what its single warp pays for (instruction fetch).
static __device__ __noinline__ float2 chroma_norm(const float r0, const float r1, const float mx, const float inv_s) {
    const bool small = mx * inv_s < FLT_MIN;
    return small ? make_float2(r0 * inv_s, r1 * inv_s) : make_float2(__fdiv_rn(r0, mx), __fdiv_rn(r1, mx));
}

// ------------------------------------------------------------------------------------------------ chroma of <= 32 frames
struct ChromaLane {
    const uint4* frag;                           // this lane's A fragments of the tuning's bank: + 128 per 32-bin step,
                                                 // + 0 / 32 / 64 / 96 = (half step 0 hi, lo, half step 1 hi, lo)
    float wny0, wny1;                            // Nyquist-bin weights of chroma g, g+8
    int g, t4;
};

// NT 8-frame tiles starting at frame f0: 32 steps of 32 bins, the next step's fragments are loaded before the current
// step's MMAs are issued; then per-frame inf-norm and the float64 running sums of this lane's (chroma, frame) entries.
template <int NT>
static __device__ __forceinline__ void chroma_group(const ChromaLane& cl, const StreamSlice& sl, const int f0, const int T,
                                                    double& cs0, double& cs1) {
    const uint4* prow[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        const int f = min(f0 + 8 * j + cl.g, T - 1);                // rows past the clip repeat its last frame, never used
        prow[j] = reinterpret_cast<const uint4*>(sl.gP16 + static_cast<size_t>(f) * kP16Row + 8 * cl.t4);
    }
    float acc[NT][4], acl[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) { acc[j][q] = 0.f; acl[j][q] = 0.f; }
    uint4 h0 = tail_ldc(cl.frag), l0 = tail_ldc(cl.frag + 32), h1 = tail_ldc(cl.frag + 64), l1 = tail_ldc(cl.frag + 96);
    uint4 pv[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j) pv[j] = tail_ld(prow[j]);
#pragma unroll 2
    for (int s = 0; s < 32; ++s) {                                  // 32 bins per step; uint4 index = 4 * s (32 halves)
        const int sn = min(s + 1, 31);
        const uint4* fn = cl.frag + sn * 128;
        const uint4 nh0 = tail_ldc(fn), nl0 = tail_ldc(fn + 32), nh1 = tail_ldc(fn + 64), nl1 = tail_ldc(fn + 96);
        uint4 npv[NT];
#pragma unroll
        for (int j = 0; j < NT; ++j) npv[j] = tail_ld(prow[j] + sn * 4);
#pragma unroll
        for (int j = 0; j < NT; ++j) {                              // half step 0: bins 8*t4 .. +3 of the step
            mma_f16(acc[j], h0.x, h0.y, h0.z, h0.w, pv[j].x, pv[j].y);
            mma_f16(acl[j], l0.x, l0.y, l0.z, l0.w, pv[j].x, pv[j].y);
        }
#pragma unroll
        for (int j = 0; j < NT; ++j) {                              // half step 1: bins 8*t4 + 4 .. +7
            mma_f16(acc[j], h1.x, h1.y, h1.z, h1.w, pv[j].z, pv[j].w);
            mma_f16(acl[j], l1.x, l1.y, l1.z, l1.w, pv[j].z, pv[j].w);
        }
        h0 = nh0; h1 = nh1; l0 = nl0; l1 = nl1;
#pragma unroll
        for (int j = 0; j < NT; ++j) pv[j] = npv[j];
    }
    constexpr float kLo = 1.0f / 2048.0f;
    const bool hi_row = cl.g < 4;                                   // chroma g+8 exists
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        const int fa = f0 + 8 * j + 2 * cl.t4;                      // this lane's two frames of tile j
        const int fc0 = min(fa, T - 1), fc1 = min(fa + 1, T - 1);
        const float2 ns0 = *reinterpret_cast<const float2*>(sl.gFv + static_cast<size_t>(fc0) * kFvStride);   // (Ny, 1/scale)
        const float2 ns1 = *reinterpret_cast<const float2*>(sl.gFv + static_cast<size_t>(fc1) * kFvStride);
        const float pn0 = ns0.x, pn1 = ns1.x, is0 = ns0.y, is1 = ns1.y;
        const float r00 = fmaf(cl.wny0, pn0, fmaf(acl[j][0], kLo, acc[j][0]));      // chroma g,   frame fa
        const float r01 = fmaf(cl.wny0, pn1, fmaf(acl[j][1], kLo, acc[j][1]));      // chroma g,   frame fa+1
        const float r10 = fmaf(cl.wny1, pn0, fmaf(acl[j][2], kLo, acc[j][2]));      // chroma g+8, frame fa
        const float r11 = fmaf(cl.wny1, pn1, fmaf(acl[j][3], kLo, acc[j][3]));      // chroma g+8, frame fa+1
        float m0 = fmaxf(fabsf(r00), hi_row ? fabsf(r10) : 0.0f);
        float m1 = fmaxf(fabsf(r01), hi_row ? fabsf(r11) : 0.0f);
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
            m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, o));
            m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, o));
        }
        if (fa < T) {
            const float2 v = chroma_norm(r00, r10, m0, is0);
            cs0 += static_cast<double>(v.x);
            if (hi_row) cs1 += static_cast<double>(v.y);
        }
        if (fa + 1 < T) {
            const float2 v = chroma_norm(r01, r11, m1, is1);
            cs0 += static_cast<double>(v.x);
            if (hi_row) cs1 += static_cast<double>(v.y);
        }
    }
}

// ------------------------------------------------------------------------------------------------ tail of one clip, one warp
// Phases 2-3 + the pooled row (the arithmetic of clip_tail in sfx_phases.cuh, reorganised for 32 threads and no barrier).
template <bool kDebug>
static __device__ __noinline__ void clip_tail_warp(const Params& p, unsigned char* slot_base, float* tile,
                                                   const double* s_edges, const int clip, const int T, const int np,
                                                   float* __restrict__ out, const int lane) {
    const DevTables& tb = p.tb;
    const StreamSlice sl = stream_slice(slot_base, p.Tmax, p.max_pk);
    // The slice was written tens of microseconds ago by the frame warps and has mostly left L2 since; one warp reading it
    // back is bound by the round trip of each load.  Bulk L2 prefetches (one instruction per range) start the DRAM reads
    // of everything the tail will stream well before it gets there: the peak records and the log-mel rows now, the
    // FP16 |X|^2 rows one 32-frame group ahead of the chroma loop.
    if (lane == 0) {
        if (np > 0) {
            for (int off = 0; off < np; off += 4096)
                bulk_prefetch_l2(sl.gRec + off, static_cast<unsigned>(min(4096, np - off)) * 16u);
        }
        for (int t0 = 0; t0 < T; t0 += 64) bulk_prefetch_l2(sl.gL + static_cast<size_t>(t0) * kMels, static_cast<unsigned>(min(64, T - t0)) * kMels * 4u);
        bulk_prefetch_l2(sl.gP16, static_cast<unsigned>(min(32, T)) * kP16Row * 2u);
    }
#ifdef SFX_STREAM_DIAG
    long long tp0 = clock64();
#define PHASE_MARK(k) do { const long long tp1 = clock64(); PROF_ADD(k, tp1 - tp0); tp0 = tp1; } while (0)
#else
#define PHASE_MARK(k) do { } while (0)
#endif
    int* hist = reinterpret_cast<int*>(tile + kTHist);
    uint2* redo_list = reinterpret_cast<uint2*>(tile + kTRedo);

    // ---- per-clip sums of the per-frame descriptors, in frame order (fixed lane -> frame assignment)
    double sum_c = 0.0, sum_r = 0.0;
    long long sum_z = 0;
    float gmx = -FLT_MAX;
    for (int t = lane; t < T; t += 32) {
        const float4 a = *reinterpret_cast<const float4*>(sl.gFv + static_cast<size_t>(t) * kFvStride);        // Ny, 1/s, E, cent
        const float4 b = *reinterpret_cast<const float4*>(sl.gFv + static_cast<size_t>(t) * kFvStride + 4);    // roll, lmax, zc, -
        sum_c += static_cast<double>(a.w);
        sum_r += static_cast<double>(b.x);
        sum_z += __float_as_int(b.z);
        gmx = fmaxf(gmx, b.y);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum_c += __shfl_xor_sync(0xffffffffu, sum_c, o);
        sum_r += __shfl_xor_sync(0xffffffffu, sum_r, o);
        sum_z += __shfl_xor_sync(0xffffffffu, sum_z, o);
    }
    gmx = warp_max(gmx);

    // ===================================== phase 2: tuning =====================================
    int tuning_idx = kTunings / 2;
    float thr = 0.0f;
    int nsel = 0, ndiff = 0;
    if (np > 0 && !(SFX_TAIL_SKIP & 1)) {
        const bool in_smem = np <= kTKeyCap;
        unsigned* keys = in_smem ? reinterpret_cast<unsigned*>(tile + kTKeys) : sl.gKey;
        unsigned char* bins = in_smem ? reinterpret_cast<unsigned char*>(tile + kTKeys + kTKeyCap) : sl.gBin;
        auto fetch = [&](int i) -> float4 {
            return i < np ? tail_ld(sl.gRec + i) : make_float4(0.f, 1.f, 1.f, __int_as_float(64));   // stand-in past the end
        };
        unsigned kor = 0u, kand = 0xffffffffu;
        int nredo = 0;
        float4 nxt[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) nxt[u] = fetch(lane + u * 32);
        for (int ib = 0; ib < np; ib += 128) {                 // warp-uniform trip count: the body holds a ballot
            const int i0 = ib + lane;
            float4 recs[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                recs[u] = nxt[u];
                nxt[u] = fetch(i0 + (4 + u) * 32);
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
                const unsigned qlo = static_cast<unsigned>(__double2loint(q)) & 0x1fffffffu;
                const unsigned qe = (static_cast<unsigned>(__double2hiint(q)) >> 20) & 0x7ffu;
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
                const unsigned pb = __float_as_uint(pitch);
                const float mant = __uint_as_float((pb & 0x007fffffu) | 0x3f800000u);
                const float w = fmaf(12.0f, lg2_approx(mant), -9.37631656229592f);
                float res = w - floorf(w);
                if (res >= 0.5f) res -= 1.0f;
                const float uf = fmaf(res, 100.0f, 50.0f);
                const float fl = floorf(uf);
                const float fr = uf - fl;
                int bin = max(0, min(kTunings - 1, static_cast<int>(fl)));
                const int i = i0 + u * 32;
                const bool sure = (fr > 0.0025f) & (fr < 0.9975f) & ((pb - 0x00800000u) < 0x7f000000u);
                const bool unsure = !sure && i < np;
                // edge cases: compacted into the redo list (all lanes redo them together after the loop); overflow inline
                const unsigned ub = __ballot_sync(0xffffffffu, unsure);
                if (ub) {
                    const int slot = nredo + __popc(ub & ((1u << lane) - 1u));
                    if (unsure) {
                        if (slot < kTRedoCap) redo_list[slot] = make_uint2(static_cast<unsigned>(i), pb);
                        else bin = peak_bin_exact(pitch, s_edges);
                    }
                    nredo += __popc(ub);
                }
                if (i < np) {
                    kor |= key;
                    kand &= key;
                    keys[i] = key;
                    bins[i] = static_cast<unsigned char>(bin);
                }
            }
        }
        kor = __reduce_or_sync(0xffffffffu, kor);
        kand = __reduce_and_sync(0xffffffffu, kand);
        __syncwarp();
        nredo = min(nredo, kTRedoCap);
        for (int j = lane; j < nredo; j += 32) {
            const uint2 e = redo_list[j];
            bins[e.x] = static_cast<unsigned char>(peak_bin_exact(__uint_as_float(e.y), s_edges));
        }
        __syncwarp();
        PHASE_MARK(6);
        PROF_ADD(11, np);
        if (kDebug) {
            // every peak again with the reference forms only; keys / bins must be identical
            for (int i = lane; i < np; i += 32) {
                const float4 rc = sl.gRec[i];
                const float sh = peak_shift_exact(rc.x, rc.y, rc.z);
                const float avg = __fsub_rn(rc.z, rc.x) * 0.5f;
                const unsigned key = fkey(__fadd_rn(rc.y, __fmul_rn(__fmul_rn(0.5f, avg), sh)));
                const double pitch_d = (static_cast<double>(__float_as_int(rc.w)) + static_cast<double>(sh)) *
                                       static_cast<double>(tb.sr) / static_cast<double>(kNfft);
                const int be = peak_bin_exact(static_cast<float>(pitch_d), s_edges);
                ndiff += (key != keys[i]) || (be != static_cast<int>(bins[i]));
            }
            ndiff = warp_sum_i(ndiff);
        }
        // ---- median of the peak magnitudes (numpy: mean of the two middle values for even counts)
        int cle = 0;
        const unsigned ka = radix_select_warp(keys, np, (np - 1) >> 1, hist, cle, kor, kand, lane);
        unsigned kb = ka;
        if ((np & 1) == 0 && cle <= (np >> 1)) {
            unsigned best = 0xffffffffu;                       // upper median = smallest key above ka
            for (int i = lane; i < np; i += 32) {
                const unsigned key = keys[i];
                if (key > ka && key < best) best = key;
            }
            kb = __reduce_min_sync(0xffffffffu, best);
        }
        const float fa = fkey_inv(ka), fb = fkey_inv(kb);
        thr = ((np & 1) == 0) ? __fmul_rn(__fadd_rn(fa, fb), 0.5f) : fa;
        const unsigned kthr = fkey(thr);
        // ---- histogram of the residual bins of peaks with mag >= median; first arg-max
#pragma unroll
        for (int q = 0; q < 4; ++q) hist[lane + 32 * q] = 0;
        __syncwarp();
        for (int ib = 0; ib < np; ib += 128) {
            const int i0 = ib + lane;
            unsigned k[4];
            int b[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool v = i0 + 32 * u < np;
                k[u] = v ? keys[i0 + 32 * u] : 0u;
                b[u] = v ? bins[i0 + 32 * u] : 0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) hist_add(hist, i0 + 32 * u < np && k[u] >= kthr, static_cast<unsigned>(b[u]), lane);
        }
        __syncwarp();
        {
            int bc = -1, bi = 1 << 20, tot = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int b = lane * 4 + q;
                if (b < kTunings) {
                    const int c = hist[b];
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
        __syncwarp();
    }
    if (kDebug) {
        if (p.dbg.clip_info && lane == 0) {
            float* ci = p.dbg.clip_info + static_cast<size_t>(clip) * 8;
            ci[0] = static_cast<float>(s_edges[tuning_idx]);
            ci[1] = gmx; ci[2] = static_cast<float>(np); ci[3] = thr;
            ci[4] = static_cast<float>(nsel); ci[5] = static_cast<float>(T);
            ci[6] = static_cast<float>(ndiff); ci[7] = 0.f;
        }
    }

    PHASE_MARK(7);
    // ===================================== phase 3a: MFCC ======================================
    // frame mean of max(logmel, gmax - 80) in float64 (even and odd frames summed separately, then added: the order of the
    // fused kernel), then the DCT once per clip.  Lane owns bands 4*lane .. 4*lane+3.
    if (!(SFX_TAIL_SKIP & 2)) {
        const float clampv = __fsub_rn(gmx, 80.0f);
        double ae[4] = {0.0, 0.0, 0.0, 0.0}, ao[4] = {0.0, 0.0, 0.0, 0.0};
        const float4* rows = reinterpret_cast<const float4*>(sl.gL) + lane;
        int t = 0;
        for (; t + 8 <= T; t += 8) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = tail_ld(rows + static_cast<size_t>(t + u) * (kMels / 4));
#pragma unroll
            for (int u = 0; u < 8; u += 2) {
                ae[0] += static_cast<double>(fmaxf(v[u].x, clampv));
                ae[1] += static_cast<double>(fmaxf(v[u].y, clampv));
                ae[2] += static_cast<double>(fmaxf(v[u].z, clampv));
                ae[3] += static_cast<double>(fmaxf(v[u].w, clampv));
                ao[0] += static_cast<double>(fmaxf(v[u + 1].x, clampv));
                ao[1] += static_cast<double>(fmaxf(v[u + 1].y, clampv));
                ao[2] += static_cast<double>(fmaxf(v[u + 1].z, clampv));
                ao[3] += static_cast<double>(fmaxf(v[u + 1].w, clampv));
            }
        }
        for (; t < T; ++t) {
            const float4 v = rows[static_cast<size_t>(t) * (kMels / 4)];
            double* a = (t & 1) ? ao : ae;
            a[0] += static_cast<double>(fmaxf(v.x, clampv));
            a[1] += static_cast<double>(fmaxf(v.y, clampv));
            a[2] += static_cast<double>(fmaxf(v.z, clampv));
            a[3] += static_cast<double>(fmaxf(v.w, clampv));
        }
        double* pool = reinterpret_cast<double*>(tile + kTHist);        // [128], 8-byte aligned (tile is 16-byte aligned)
#pragma unroll
        for (int j = 0; j < 4; ++j) pool[4 * lane + j] = (ae[j] + ao[j]) / static_cast<double>(T);
        __syncwarp();
        for (int k0 = 0; k0 < p.n_mfcc; k0 += 32) {
            const int k = k0 + lane;
            if (k < p.n_mfcc) {
                double d0 = 0.0, d1 = 0.0;                   // bands 0..63 and 64..127, added in that order (as clip_tail)
#pragma unroll 8
                for (int q = 0; q < 64; ++q) {
                    d0 = fma(tb.dctT[q * kMels + k], pool[q], d0);
                    d1 = fma(tb.dctT[(q + 64) * kMels + k], pool[q + 64], d1);
                }
                out[k] = static_cast<float>(d0 + d1);
            }
        }
        __syncwarp();
    }

    PHASE_MARK(8);
    // ===================================== phase 3b: chroma ====================================
    // raw[c][f] = sum_k W[c][k] |X|^2[k][f]: m16n8k16 FP16 MMAs, A = bank rows (hi and 2^11*lo, separate accumulators) read
    // from L2, B = the frames' scaled FP16 |X|^2 rows.  Lane (g = lane/4, t4 = lane%4) feeds bank rows g, g+8 and frame g of
    // each of the group's 4 tiles with its 8 contiguous bins of every 32-bin step (the same K permutation on both
    // operands leaves the products unchanged).  D fragment: [0..1] = (chroma g, frames 2*t4, 2*t4+1 of the tile),
    // [2..3] = (chroma g+8).
    if (!(SFX_TAIL_SKIP & 4)) {
        const int g = lane >> 2, t4 = lane & 3;
        const float wny0 = __ldg(tb.chroma_ny + tuning_idx * kChroma + g);
        const float wny1 = g < 4 ? __ldg(tb.chroma_ny + tuning_idx * kChroma + g + 8) : 0.0f;
        double cs0 = 0.0, cs1 = 0.0;                                // sums over this lane's frames of chroma g / g+8
        const ChromaLane cl{tb.chroma_frag + static_cast<size_t>(tuning_idx) * (32 * 128) + lane, wny0, wny1, g, t4};
        for (int f0 = 0; f0 < T; f0 += 32) {
            if (lane == 0 && f0 + 32 < T)                           // next group's rows on their way to L2
                bulk_prefetch_l2(sl.gP16 + static_cast<size_t>(f0 + 32) * kP16Row, static_cast<unsigned>(min(32, T - f0 - 32)) * kP16Row * 2u);
            // always four 8-frame tiles (one instantiation: code size): tiles past the clip repeat its last frame, are
            // multiplied for nothing (the tensor pipe is otherwise idle) and ignored
            chroma_group<4>(cl, sl, f0, T, cs0, cs1);
        }
        cs0 += __shfl_xor_sync(0xffffffffu, cs0, 1);
        cs1 += __shfl_xor_sync(0xffffffffu, cs1, 1);
        cs0 += __shfl_xor_sync(0xffffffffu, cs0, 2);
        cs1 += __shfl_xor_sync(0xffffffffu, cs1, 2);
        const double invT = 1.0 / static_cast<double>(T);
        if (t4 == 0) {
            out[p.n_mfcc + g] = static_cast<float>(cs0 * invT);
            if (g < 4) out[p.n_mfcc + 8 + g] = static_cast<float>(cs1 * invT);
        }
    }

    PHASE_MARK(9);
    // ===================================== epilogue: pooled descriptors ========================
    {
        // pooled rms from the hop energies (frame t spans hops t-2 .. t+1; hops outside [0, T) are zero padding)
        double a = 0.0;
        for (int t = lane; t < T; t += 32) {
            const float* gE = sl.gFv + static_cast<size_t>(t) * kFvStride + 2;      // hop energies, kFvStride apart
            float e = (t >= 2) ? gE[-2 * kFvStride] : 0.0f;
            e += (t >= 1) ? gE[-kFvStride] : 0.0f;
            e += gE[0];
            e += (t + 1 < T) ? gE[kFvStride] : 0.0f;
            const float r = sqrtf(e * (1.0f / kNfft));
            a += static_cast<double>(r);
            if (kDebug) {
                if (p.dbg.frame_feat && t < p.dbg.T_dbg)
                    p.dbg.frame_feat[(static_cast<size_t>(clip) * p.dbg.T_dbg + t) * 4 + 2] = r;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) {
            const double invT = 1.0 / static_cast<double>(T);
            out[p.n_mfcc + 12] = static_cast<float>(static_cast<double>(sum_z) / (static_cast<double>(kNfft) * T));
            out[p.n_mfcc + 13] = static_cast<float>(sum_c * invT);
            out[p.n_mfcc + 14] = static_cast<float>(sum_r * invT);
            out[p.n_mfcc + 15] = static_cast<float>(a / static_cast<double>(T));
        }
    }
    __syncwarp();
    PHASE_MARK(10);
}

// ------------------------------------------------------------------------------------------------ scheduler
// Next item for the calling warp (lane 0 only).  Returns the kind; slot / t / clip through the references.
//   1. a clip whose frames are all done: its tail comes first (it frees a slot); claimed by clearing its ready bit
//   2. a frame: one 64-bit atomicAdd on the ticket word
//   3. the ticket is exhausted: one warp (spin lock `opener`, taken once per clip) pulls the next clip of the batch into
//      a free slot and publishes its ticket; it keeps frame 0 for itself
//   4. nothing to hand out: exit when the batch is exhausted and every slot is free, else poll again
__device__ __forceinline__ int next_work(const Params& p, Sched* sc, int* counter, int& slot, int& t, int& clip) {
    volatile Sched* v = sc;
    const int ready = v->ready_mask;
    if (ready) {
        const int s = __ffs(ready) - 1;
        if (atomicAnd(&sc->ready_mask, ~(1 << s)) & (1 << s)) {
            __threadfence_block();                       // acquire: the frames' rows and records
            slot = s;
            return kWorkTail;
        }
        return kWorkWait;                                // somebody else took it: look again
    }
    {
        unsigned h = v->hot;                              // look first: polls of an exhausted ticket must not advance it
        if (static_cast<int>(h & 0xffffffu) < v->ring_T[(h >> 24) & 7u]) {
            h = atomicAdd(&sc->hot, 1u);
            const int g = (h >> 24) & 7u, tt = static_cast<int>(h & 0xffffffu);
            if (tt < v->ring_T[g]) {
                slot = v->ring_slot[g];
                t = tt;
                return kWorkFrame;
            }
        }
    }
    if (v->qdone) return v->active == 0 ? kWorkExit : kWorkWait;
    if (atomicCAS(&sc->opener, 0, 1) != 0) return kWorkWait;
    __threadfence_block();
    int kind = kWorkWait;
    const unsigned h2 = v->hot;                          // a clip may have been published while we queued for the lock
    const bool exhausted = static_cast<int>(h2 & 0xffffffu) >= v->ring_T[(h2 >> 24) & 7u];
    const int fm = v->free_mask;
    if (exhausted && fm && !v->qdone) {
        const int q = atomicAdd(counter, 1);
        if (q >= p.B) {
            v->qdone = 1;
        } else {
            clip = p.order ? p.order[q] : q;
            const long long n = clip_samples(p, clip);
            if (n <= 0) {
                kind = kWorkBad;
            } else {
                const int s = __ffs(fm) - 1;
                const int T = 1 + static_cast<int>(n / kHop);
                atomicAnd(&sc->free_mask, ~(1 << s));
                atomicAdd(&sc->active, 1);
                v->clip[s] = clip; v->n[s] = n; v->T[s] = T; v->done[s] = 0; v->npk[s] = 0;
                const unsigned gen = ((h2 >> 24) + 1u) & 0xffu;
                v->ring_slot[gen & 7u] = s; v->ring_T[gen & 7u] = T;
                __threadfence_block();
                atomicExch(&sc->hot, (gen << 24) | 1u);          // frame 0 is ours
                slot = s; t = 0; kind = kWorkFrame;
            }
        }
    }
    __threadfence_block();
    atomicExch(&sc->opener, 0);
    return kind;
}

// ------------------------------------------------------------------------------------------------ kernel
template <bool kDebug>
__global__ void __launch_bounds__(kSThreads, 1) sfx_stream_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* s_hann = reinterpret_cast<float2*>(smem_raw);
    float2* s_tw1 = s_hann + 64;                     // (s_hann: 32 lanes x (cos, cos', sin, sin'), fill_hann_phases)
    float2* s_tw2 = s_tw1 + 1024;
    float2* s_melab = s_tw2 + 512;                                           // [33*32]  (tw2: rows k2 < 16 only)
    float* s_ex = reinterpret_cast<float*>(s_melab + 17 * 64);               // [kSW][kExFloats]
    double* s_edges = reinterpret_cast<double*>(s_ex + kSW * kExFloats);     // [104]
    Sched* sc = reinterpret_cast<Sched*>(s_edges + 104);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const DevTables& tb = p.tb;
    fill_hann_phases(s_hann, tid);
    for (int i = tid; i < 1024; i += kSThreads) {
        const int r = i >> 5, l = i & 31;
        const int d = (r >> 1) * 64 + 2 * l + (r & 1);
        s_tw1[d] = tb.tw1[i];
        if (i < 512) s_tw2[d] = tb.tw2[i];
    }
    for (int i = tid; i < 33 * 32; i += kSThreads) {
        const int r = i >> 5, l = i & 31;
        s_melab[(r >> 1) * 64 + 2 * l + (r & 1)] = tb.mel_ab[i];
    }
    for (int i = tid; i <= kTunings; i += kSThreads) s_edges[i] = tb.edges[i];
    for (int i = tid; i < static_cast<int>(sizeof(Sched) / 4); i += kSThreads) reinterpret_cast<int*>(sc)[i] = 0;
    __syncthreads();
    if (tid == 0) sc->free_mask = (1 << kSlots) - 1;        // hot = generation 0, ring_T[0] = 0: no frames yet
    __syncthreads();

    FrameSmem fs;
    fs.s_hann = s_hann; fs.s_tw1 = s_tw1; fs.s_tw2 = s_tw2; fs.s_melab = s_melab;
    fs.Pb = s_ex + warp * kExFloats;
    fs.ex = reinterpret_cast<float2*>(fs.Pb);
    fs.part = fs.Pb + kPartOff;
    fs.mel_mask = tb.mel_mask[lane]; fs.mel_ps = tb.mel_ps;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const int* q = tb.mel_src + (32 * s + lane) * 3;
        fs.msrc[s] = q[0] | (q[1] << 10) | (q[2] << 20);
    }
    fs.s_msrc = nullptr;
    fs.bin_hz = static_cast<float>(static_cast<double>(tb.sr) / kNfft);
    fs.aligned8 = p.aligned8 != 0;

    int* counter = reinterpret_cast<int*>(p.ws);
    unsigned char* slots_base = p.ws + kWsHeader + static_cast<size_t>(blockIdx.x) * kSlots * p.cta_scratch_bytes;
    volatile Sched* v = sc;

    long long wait_since = 0;                            // clock of the first of a run of empty-handed polls
    for (;;) {
        int kind = 0, slot = 0, t = 0, clip = 0;
#ifdef SFX_STREAM_DIAG
        const long long tk0 = clock64();
#endif
        if (lane == 0) kind = next_work(p, sc, counter, slot, t, clip);
        kind = __shfl_sync(0xffffffffu, kind, 0);
#ifdef SFX_STREAM_DIAG
        const long long tk1 = clock64();
        PROF_ADD(3, tk1 - tk0);
#endif
        if (kind == kWorkExit) break;
        if (kind == kWorkWait) {
            // nothing to hand out right now (every slot is waiting for straggler frames or in its tail)
            const long long now = clock64();
            if (wait_since == 0) wait_since = now;
#ifdef SFX_STREAM_DIAG
            if (__shfl_sync(0xffffffffu, now - wait_since > 1000000000ll ? 1 : 0, 0)) { if (lane == 0) diag_dump(sc, warp, 1); break; }
#endif
            if (now - wait_since > 16000000000ll) __trap();     // ~8 s: a scheduler bug must fail loudly, not hang the GPU
            __nanosleep(256);
#ifdef SFX_STREAM_DIAG
            PROF_ADD(2, clock64() - tk1);
#endif
            continue;
        }
        wait_since = 0;
        slot = __shfl_sync(0xffffffffu, slot, 0);
        t = __shfl_sync(0xffffffffu, t, 0);
        clip = __shfl_sync(0xffffffffu, clip, 0);
        if (kind == kWorkBad) {                          // length <= 0 or beyond the scratch slice: a row of NaN
            float* out = p.out + static_cast<long long>(clip) * p.out_stride;
            for (int i = lane; i < p.n_mfcc + 16; i += 32) out[i] = __int_as_float(0x7fc00000);
            continue;
        }
        unsigned char* slot_base = slots_base + static_cast<size_t>(slot) * p.cta_scratch_bytes;
        if (kind == kWorkFrame) {
            const StreamSlice sl = stream_slice(slot_base, p.Tmax, p.max_pk);
            clip = v->clip[slot];
            const long long n = v->n[slot];
            const int T = v->T[slot];
            const float* x = p.wave + static_cast<long long>(clip) * p.row_stride;
            FrameOut fo;
            fo.gP16 = sl.gP16; fo.gL = sl.gL; fo.gRec = sl.gRec; fo.gE = nullptr; fo.gNy = nullptr; fo.gInvS = nullptr;
            fo.npk = nullptr; fo.gSeg = nullptr; fo.s_wacc = nullptr; fo.s_f = nullptr; fo.cursor = &sc->npk[slot];
            fo.s_lm = nullptr; fo.s_lmin = nullptr;
            fo.gCent = nullptr; fo.gRoll = nullptr; fo.gLmax = nullptr; fo.gZc = nullptr; fo.gFv = sl.gFv;
            int unused_zc = 0, wcount = 0;
            process_frame<kDebug, kModeStream>(p, tb, fs, fo, x, n, T, t, clip, lane, warp, unused_zc, wcount);
            if (lane == 0) {
                __threadfence_block();                   // the frame's rows and records before the completion count
                if (atomicAdd(&sc->done[slot], 1) + 1 == T) atomicOr(&sc->ready_mask, 1 << slot);
            }
            __syncwarp();
#ifdef SFX_STREAM_DIAG
            PROF_ADD(0, clock64() - tk1);
            PROF_ADD(4, 1);
#endif
        } else {                                         // kWorkTail
            __threadfence_block();
            clip = v->clip[slot];
            const int T = v->T[slot];
            float* out = p.out + static_cast<long long>(clip) * p.out_stride;
#ifndef SFX_STREAM_NOTAIL                                // (experiment: frames only, rows are garbage)
            clip_tail_warp<kDebug>(p, slot_base, fs.Pb, s_edges, clip, T, v->npk[slot], out, lane);
#endif
            if (lane == 0) {
                __threadfence_block();
                atomicOr(&sc->free_mask, 1 << slot);
                atomicSub(&sc->active, 1);
            }
            __syncwarp();
#ifdef SFX_STREAM_DIAG
            PROF_ADD(1, clock64() - tk1);
            PROF_ADD(5, 1);
#endif
        }
    }
}

// ------------------------------------------------------------------------------------------------ host
size_t smem_stream() {
    return sizeof(float2) * (64 + 1536 + 17 * 64) + sizeof(float) * kSW * kExFloats + sizeof(double) * 104 + sizeof(Sched);
}
int stream_slots() { return kSlots; }

cudaError_t configure_stream(int* blocks_per_sm) {
    const int smem = static_cast<int>(smem_stream());
    cudaError_t e = cudaFuncSetAttribute(sfx_stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(sfx_stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, sfx_stream_kernel<false>, kSThreads, smem);
}

cudaError_t launch_stream(const Params& p, int grid, bool debug, cudaStream_t stream) {
    const size_t smem = smem_stream();
    if (debug) sfx_stream_kernel<true><<<grid, kSThreads, smem, stream>>>(p);
    else       sfx_stream_kernel<false><<<grid, kSThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace sfx
