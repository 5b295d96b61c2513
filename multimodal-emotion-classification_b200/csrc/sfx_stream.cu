// sfx_stream.cu -- the third pipeline of the extractor: one persistent 16-warp CTA per SM in which the per-clip tail never
// idles the FFT warps.
//
// The fused kernel (sfx_kernels.cu) runs a clip's tail (tuning estimate, MFCC, chroma, pooled row) with all 8 warps of its
// CTA: the tail holds ~12 % of the instructions but 25-40 % of a CTA's time, because it is a chain of scratch round trips
// and CTA barriers, and it only overlaps with FFT work of the *other* co-resident CTA.  Here every warp is an independent
// worker that pulls items from a CTA-local scheduler (a few words of shared memory behind a spin lock):
//   FRAME(slot, t)  phase 1 of STFT frame t of the clip in `slot` (process_frame<., kModeStream>: the code of the other
//                   pipelines); frames of the next clip are handed out as soon as the current clip's last frame has been
//                   *taken*, so there is no per-clip barrier and no round imbalance (130 frames over 8 warps);
//   TAIL(slot)      phases 2-3 of a clip whose frames are all done, executed by ONE warp from start to end
//                   (clip_tail_warp below: no CTA barrier anywhere).  A tail is latency-bound whoever runs it; run by one
//                   warp it costs one warp's time instead of eight, and the other 15 keep transforming frames.
// A CTA owns kSlots scratch slices; a slot is FREE -> FRAMES -> READY -> TAIL -> FREE.  Per-frame descriptors are stored in
// the slice and summed by the tail warp in frame order, peak records go to the (slot, warp) segment of whichever warp ran
// the frame, so the pooled row does not depend on which warp ran which frame.
//
// Chroma in the tail: raw = W . |X|^2 on the tensor cores (m16n8k16, FP16 hi/lo bank, FP32 accumulate) with the bank read
// straight from L2 (all 100 banks are 5 MB and resident): a bank fragment is applied to 32 frames (4 MMA N-tiles) per load,
// so no shared-memory copy of the bank is needed and any number of tails can run side by side.
#include "sfx_phases.cuh"

#ifndef SFX_STREAM_SLOTS
#define SFX_STREAM_SLOTS 4
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

enum : int { kFree = 0, kFrames = 1, kReady = 2, kTail = 3 };
enum : int { kWorkExit = 0, kWorkWait = 1, kWorkFrame = 2, kWorkTail = 3, kWorkBad = 4 };

struct Sched {                       // shared memory, every field accessed through volatile or atomics
    int lock, qdone;
    int state[kSlots], clip[kSlots], T[kSlots], next[kSlots], done[kSlots];
    long long n[kSlots];
    int cnt[kSlots][kSW];            // peak records in segment (slot, warp)
};

struct StreamSlice {
    __half* gP16; float* gL; float* gFv; float4* gRec; unsigned* gKey; unsigned char* gBin;
};

// layout of a slot: FP16 |X|^2 rows | log-mel rows | per-frame value records | record segments | keys | bins
__device__ __forceinline__ StreamSlice stream_slice(unsigned char* base, int Tmax, int max_pk) {
    StreamSlice s;
    s.gP16 = reinterpret_cast<__half*>(base);
    s.gL = reinterpret_cast<float*>(s.gP16 + static_cast<size_t>(Tmax) * kP16Stride);
    s.gFv = s.gL + static_cast<size_t>(Tmax) * kMels;
    s.gRec = reinterpret_cast<float4*>(s.gFv + static_cast<size_t>(Tmax) * kFvStride);
    s.gKey = reinterpret_cast<unsigned*>(s.gRec + static_cast<size_t>(kSW) * stream_seg_frames(Tmax) * max_pk);
    s.gBin = reinterpret_cast<unsigned char*>(s.gKey + static_cast<size_t>(Tmax) * max_pk);
    return s;
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
        for (int i0 = lane; i0 < np; i0 += 128) {
            unsigned k[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) k[u] = (i0 + 32 * u < np) ? keys[i0 + 32 * u] : 0u;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i0 + 32 * u < np && (k[u] & mask) == prefix) atomicAdd(&hist[(k[u] >> shift) & bmask], 1);
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

// ------------------------------------------------------------------------------------------------ chroma of <= 32 frames
struct ChromaLane {
    const uint4 *whi0, *whi1, *wlo0, *wlo1;      // this lane's bank rows g, g+8 (hi and lo halves) at its 8 bins of a step
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
        prow[j] = reinterpret_cast<const uint4*>(sl.gP16 + static_cast<size_t>(f) * kP16Stride + 8 * cl.t4);
    }
    float acc[NT][4], acl[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) { acc[j][q] = 0.f; acl[j][q] = 0.f; }
    uint4 h0 = __ldg(cl.whi0), h1 = __ldg(cl.whi1), l0 = __ldg(cl.wlo0), l1 = __ldg(cl.wlo1);
    uint4 pv[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j) pv[j] = prow[j][0];
#pragma unroll 2
    for (int s = 0; s < 32; ++s) {                                  // 32 bins per step; uint4 index = 4 * s (32 halves)
        const int sn = min(s + 1, 31) * 4;
        const uint4 nh0 = __ldg(cl.whi0 + sn), nh1 = __ldg(cl.whi1 + sn);
        const uint4 nl0 = __ldg(cl.wlo0 + sn), nl1 = __ldg(cl.wlo1 + sn);
        uint4 npv[NT];
#pragma unroll
        for (int j = 0; j < NT; ++j) npv[j] = prow[j][sn];
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            mma_f16(acc[j], h0.x, h1.x, h0.y, h1.y, pv[j].x, pv[j].y);
            mma_f16(acl[j], l0.x, l1.x, l0.y, l1.y, pv[j].x, pv[j].y);
        }
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            mma_f16(acc[j], h0.z, h1.z, h0.w, h1.w, pv[j].z, pv[j].w);
            mma_f16(acl[j], l0.z, l1.z, l0.w, l1.w, pv[j].z, pv[j].w);
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
        const float2 ns0 = *reinterpret_cast<const float2*>(sl.gFv + static_cast<size_t>(fc0) * kFvStride + 1);   // (Ny, 1/scale)
        const float2 ns1 = *reinterpret_cast<const float2*>(sl.gFv + static_cast<size_t>(fc1) * kFvStride + 1);
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
        // librosa.util.normalize: lengths below tiny(float32) are replaced by 1 (in unscaled units)
        if (fa < T) {
            const bool small = m0 * is0 < FLT_MIN;
            cs0 += static_cast<double>(small ? r00 * is0 : __fdiv_rn(r00, m0));
            if (hi_row) cs1 += static_cast<double>(small ? r10 * is0 : __fdiv_rn(r10, m0));
        }
        if (fa + 1 < T) {
            const bool small = m1 * is1 < FLT_MIN;
            cs0 += static_cast<double>(small ? r01 * is1 : __fdiv_rn(r01, m1));
            if (hi_row) cs1 += static_cast<double>(small ? r11 * is1 : __fdiv_rn(r11, m1));
        }
    }
}

// ------------------------------------------------------------------------------------------------ tail of one clip, one warp
// Phases 2-3 + the pooled row (the arithmetic of clip_tail in sfx_phases.cuh, reorganised for 32 threads and no barrier).
template <bool kDebug>
static __device__ __noinline__ void clip_tail_warp(const Params& p, unsigned char* slot_base, const volatile int* seg_cnt,
                                                   const int seg_cap, float* tile, const double* s_edges, const int clip,
                                                   const int T, float* __restrict__ out, const int lane) {
    const DevTables& tb = p.tb;
    const StreamSlice sl = stream_slice(slot_base, p.Tmax, p.max_pk);
    int* hist = reinterpret_cast<int*>(tile + kTHist);
    uint2* redo_list = reinterpret_cast<uint2*>(tile + kTRedo);

    // ---- per-clip sums of the per-frame descriptors, in frame order (fixed lane -> frame assignment)
    double sum_c = 0.0, sum_r = 0.0;
    long long sum_z = 0;
    float gmx = -FLT_MAX;
    for (int t = lane; t < T; t += 32) {
        const float4 a = *reinterpret_cast<const float4*>(sl.gFv + static_cast<size_t>(t) * kFvStride);        // E, Ny, 1/s, cent
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
    int np = 0;
#pragma unroll
    for (int w = 0; w < kSW; ++w) np += seg_cnt[w];
    int tuning_idx = kTunings / 2;
    float thr = 0.0f;
    int nsel = 0, ndiff = 0;
    if (np > 0) {
        const bool in_smem = np <= kTKeyCap;
        unsigned* keys = in_smem ? reinterpret_cast<unsigned*>(tile + kTKeys) : sl.gKey;
        unsigned char* bins = in_smem ? reinterpret_cast<unsigned char*>(tile + kTKeys + kTKeyCap) : sl.gBin;
        // dense peak index -> record: indices only grow, so the segment boundaries are walked once
        int seg_w = 0, seg_end = seg_cnt[0], seg_adj = 0;
        auto fetch = [&](int i) -> float4 {
            if (i >= np) return make_float4(0.f, 1.f, 1.f, __int_as_float(64));       // harmless stand-in past the end
            while (i >= seg_end) {
                seg_adj += seg_cap - seg_cnt[seg_w];
                ++seg_w;
                seg_end += seg_cnt[seg_w];
            }
            return sl.gRec[i + seg_adj];
        };
        unsigned kor = 0u, kand = 0xffffffffu;
        int nredo = 0;
        float4 nxt[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) nxt[u] = fetch(lane + u * 32);
        for (int i0 = lane; i0 < np; i0 += 128) {
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
        if (kDebug) {
            // every peak again with the reference forms only; keys / bins must be identical
            int sw = 0, send = seg_cnt[0], sadj = 0;
            for (int i = lane; i < np; i += 32) {
                while (i >= send) {
                    sadj += seg_cap - seg_cnt[sw];
                    ++sw;
                    send += seg_cnt[sw];
                }
                const float4 rc = sl.gRec[i + sadj];
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
        for (int i0 = lane; i0 < np; i0 += 128) {
            unsigned k[4];
            int b[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool v = i0 + 32 * u < np;
                k[u] = v ? keys[i0 + 32 * u] : 0u;
                b[u] = v ? bins[i0 + 32 * u] : 0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i0 + 32 * u < np && k[u] >= kthr) atomicAdd(&hist[b[u]], 1);
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

    // ===================================== phase 3a: MFCC ======================================
    // frame mean of max(logmel, gmax - 80) in float64 (even and odd frames summed separately, then added: the order of the
    // fused kernel), then the DCT once per clip.  Lane owns bands 4*lane .. 4*lane+3.
    {
        const float clampv = __fsub_rn(gmx, 80.0f);
        double ae[4] = {0.0, 0.0, 0.0, 0.0}, ao[4] = {0.0, 0.0, 0.0, 0.0};
        const float4* rows = reinterpret_cast<const float4*>(sl.gL) + lane;
        int t = 0;
        for (; t + 8 <= T; t += 8) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = rows[static_cast<size_t>(t + u) * (kMels / 4)];
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
                double d = 0.0;
#pragma unroll 8
                for (int q = 0; q < kMels; ++q) d = fma(tb.dctT[q * kMels + k], pool[q], d);
                out[k] = static_cast<float>(d);
            }
        }
        __syncwarp();
    }

    // ===================================== phase 3b: chroma ====================================
    // raw[c][f] = sum_k W[c][k] |X|^2[k][f]: m16n8k16 FP16 MMAs, A = bank rows (hi and 2^11*lo, separate accumulators) read
    // from L2, B = the frames' scaled FP16 |X|^2 rows.  Lane (g = lane/4, t4 = lane%4) feeds bank rows g, g+8 and frame g of
    // each of the group's 4 tiles with its 8 contiguous bins of every 32-bin step (the same K permutation on both
    // operands leaves the products unchanged).  D fragment: [0..1] = (chroma g, frames 2*t4, 2*t4+1 of the tile),
    // [2..3] = (chroma g+8).
    {
        const int g = lane >> 2, t4 = lane & 3;
        const int r1 = (g < 4) ? g + 8 : g;                         // bank rows 12..15 do not exist
        const __half* bank = reinterpret_cast<const __half*>(tb.chroma16) + static_cast<size_t>(tuning_idx) * 2 * kChroma * kP16Stride;
        const uint4* whi0 = reinterpret_cast<const uint4*>(bank + g * kP16Stride + 8 * t4);
        const uint4* whi1 = reinterpret_cast<const uint4*>(bank + r1 * kP16Stride + 8 * t4);
        const uint4* wlo0 = reinterpret_cast<const uint4*>(bank + (kChroma + g) * kP16Stride + 8 * t4);
        const uint4* wlo1 = reinterpret_cast<const uint4*>(bank + (kChroma + r1) * kP16Stride + 8 * t4);
        const float wny0 = __ldg(tb.chroma_ny + tuning_idx * kChroma + g);
        const float wny1 = __ldg(tb.chroma_ny + tuning_idx * kChroma + r1);
        double cs0 = 0.0, cs1 = 0.0;                                // sums over this lane's frames of chroma g / g+8
        const ChromaLane cl{whi0, whi1, wlo0, wlo1, wny0, wny1, g, t4};
        for (int f0 = 0; f0 < T; f0 += 32) {
            const int nt = min(4, (T - f0 + 7) >> 3);               // 8-frame tiles of this group
            if (nt == 4)      chroma_group<4>(cl, sl, f0, T, cs0, cs1);
            else if (nt == 3) chroma_group<3>(cl, sl, f0, T, cs0, cs1);
            else if (nt == 2) chroma_group<2>(cl, sl, f0, T, cs0, cs1);
            else              chroma_group<1>(cl, sl, f0, T, cs0, cs1);
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

    // ===================================== epilogue: pooled descriptors ========================
    {
        // pooled rms from the hop energies (frame t spans hops t-2 .. t+1; hops outside [0, T) are zero padding)
        double a = 0.0;
        for (int t = lane; t < T; t += 32) {
            const float* gE = sl.gFv + static_cast<size_t>(t) * kFvStride;          // hop energies, kFvStride apart
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
}

// ------------------------------------------------------------------------------------------------ scheduler
__device__ __forceinline__ void sched_lock(Sched* sc) {
    const long long t0 = clock64();
    while (atomicCAS(&sc->lock, 0, 1) != 0) {
        __nanosleep(32);
        if (clock64() - t0 > 16000000000ll) __trap();           // the critical sections are a few hundred cycles long
    }
    __threadfence_block();
}
__device__ __forceinline__ void sched_unlock(Sched* sc) {
    __threadfence_block();
    atomicExch(&sc->lock, 0);
}

// Next item for warp `warp` (called by lane 0).  Returns the kind; slot / t / clip through the references.
__device__ __forceinline__ int next_work(const Params& p, Sched* sc, int* counter, const int warp, const int seg_cap,
                                         int& slot, int& t, int& clip) {
    volatile Sched* v = sc;
    int kind = kWorkWait;
    sched_lock(sc);
    // 1. a clip whose frames are all done: its tail comes first (it frees a slot)
#pragma unroll
    for (int s = 0; s < kSlots; ++s)
        if (kind == kWorkWait && v->state[s] == kReady) { v->state[s] = kTail; slot = s; kind = kWorkTail; }
    // 2. a frame of a clip that is being transformed (this warp's record segment must have room for a full frame)
#pragma unroll
    for (int s = 0; s < kSlots; ++s)
        if (kind == kWorkWait && v->state[s] == kFrames && v->next[s] < v->T[s] && v->cnt[s][warp] + p.max_pk <= seg_cap) {
            slot = s; t = v->next[s]; v->next[s] = t + 1; kind = kWorkFrame;
        }
    // 3. the next clip of the batch into a free slot
    if (kind == kWorkWait && !v->qdone) {
        int fs = -1;
#pragma unroll
        for (int s = 0; s < kSlots; ++s)
            if (fs < 0 && v->state[s] == kFree) fs = s;
        if (fs >= 0) {
            const int q = atomicAdd(counter, 1);
            if (q >= p.B) {
                v->qdone = 1;
            } else {
                clip = p.order ? p.order[q] : q;
                const long long n = clip_samples(p, clip);
                if (n <= 0) {
                    kind = kWorkBad;
                } else {
                    v->clip[fs] = clip; v->n[fs] = n; v->T[fs] = 1 + static_cast<int>(n / kHop);
                    v->next[fs] = 1; v->done[fs] = 0;
#pragma unroll
                    for (int w = 0; w < kSW; ++w) v->cnt[fs][w] = 0;
                    v->state[fs] = kFrames;
                    slot = fs; t = 0; kind = kWorkFrame;
                }
            }
        }
    }
    // 4. nothing to hand out: done when the batch is exhausted and every slot is free
    if (kind == kWorkWait && v->qdone) {
        bool all_free = true;
#pragma unroll
        for (int s = 0; s < kSlots; ++s) all_free &= v->state[s] == kFree;
        if (all_free) kind = kWorkExit;
    }
    sched_unlock(sc);
    return kind;
}

// ------------------------------------------------------------------------------------------------ kernel
template <bool kDebug>
__global__ void __launch_bounds__(kSThreads, 1) sfx_stream_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* s_hann = reinterpret_cast<float2*>(smem_raw);
    float2* s_tw1 = s_hann + 1024;
    float2* s_tw2 = s_tw1 + 1024;
    float2* s_melab = s_tw2 + 512;                                           // [33*32]  (tw2: rows k2 < 16 only)
    float* s_ex = reinterpret_cast<float*>(s_melab + 17 * 64);               // [kSW][kExFloats]
    double* s_edges = reinterpret_cast<double*>(s_ex + kSW * kExFloats);     // [104]
    Sched* sc = reinterpret_cast<Sched*>(s_edges + 104);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const DevTables& tb = p.tb;
    for (int i = tid; i < 1024; i += kSThreads) {
        const int r = i >> 5, l = i & 31;
        const int d = (r >> 1) * 64 + 2 * l + (r & 1);
        s_hann[(r & 15) * 64 + 2 * l + (r >> 4)] = tb.hann[i];        // rows r, r + 16 side by side
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
    fs.bin_hz = static_cast<float>(static_cast<double>(tb.sr) / kNfft);
    fs.aligned8 = p.aligned8 != 0;

    int* counter = reinterpret_cast<int*>(p.ws);
    unsigned char* slots_base = p.ws + kWsHeader + static_cast<size_t>(blockIdx.x) * kSlots * p.cta_scratch_bytes;
    const int seg_cap = stream_seg_frames(p.Tmax) * p.max_pk;
    volatile Sched* v = sc;

    long long wait_since = 0;                            // clock of the first of a run of empty-handed polls
    for (;;) {
        int kind = 0, slot = 0, t = 0, clip = 0;
        if (lane == 0) kind = next_work(p, sc, counter, warp, seg_cap, slot, t, clip);
        kind = __shfl_sync(0xffffffffu, kind, 0);
        if (kind == kWorkExit) break;
        if (kind == kWorkWait) {
            // nothing to hand out right now (every slot is waiting for straggler frames or in its tail)
            const long long now = clock64();
            if (wait_since == 0) wait_since = now;
            if (now - wait_since > 16000000000ll) __trap();     // ~8 s: a scheduler bug must fail loudly, not hang the GPU
            __nanosleep(256);
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
            fo.npk = nullptr; fo.gSeg = sl.gRec + static_cast<size_t>(warp) * seg_cap; fo.s_wacc = nullptr; fo.s_f = nullptr;
            fo.gCent = nullptr; fo.gRoll = nullptr; fo.gLmax = nullptr; fo.gZc = nullptr; fo.gFv = sl.gFv;
            int unused_zc = 0, wcount = v->cnt[slot][warp];
            process_frame<kDebug, kModeStream>(p, tb, fs, fo, x, n, T, t, clip, lane, warp, unused_zc, wcount);
            if (lane == 0) {
                v->cnt[slot][warp] = wcount;
                __threadfence_block();                   // the frame's rows and records before the completion count
                if (atomicAdd(&sc->done[slot], 1) + 1 == T) {
                    __threadfence_block();
                    v->state[slot] = kReady;
                }
            }
            __syncwarp();
        } else {                                         // kWorkTail
            __threadfence_block();
            clip = v->clip[slot];
            const int T = v->T[slot];
            float* out = p.out + static_cast<long long>(clip) * p.out_stride;
            clip_tail_warp<kDebug>(p, slot_base, &v->cnt[slot][0], seg_cap, fs.Pb, s_edges, clip, T, out, lane);
            if (lane == 0) {
                __threadfence_block();
                v->state[slot] = kFree;
            }
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------------ host
size_t smem_stream() {
    return sizeof(float2) * (2560 + 17 * 64) + sizeof(float) * kSW * kExFloats + sizeof(double) * 104 + sizeof(Sched);
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
