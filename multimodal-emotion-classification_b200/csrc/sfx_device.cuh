// sfx_device.cuh -- device helpers shared by the fused kernel (sfx_kernels.cu) and the split kernels (sfx_split.cu):
// compile-time unrolling, the register-resident 32-point FFT (scalar and packed-FP32 forms), warp reductions, the CTA
// radix select, MMA / TMA wrappers, small inline-PTX helpers and the frame loader.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cfloat>
#include <cstdint>
#include <type_traits>
#include <utility>

#include "sfx_internal.h"

namespace sfx {

// ------------------------------------------------------------------------------------------------
// compile-time helpers
template <int... I, class F>
__device__ __forceinline__ void sfor_impl(std::integer_sequence<int, I...>, F&& f) {
    (f(std::integral_constant<int, I>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void sfor(F&& f) {
    sfor_impl(std::make_integer_sequence<int, N>{}, static_cast<F&&>(f));
}

__host__ __device__ constexpr int brev5(int k) {
    return ((k & 1) << 4) | ((k & 2) << 2) | (k & 4) | ((k & 8) >> 2) | ((k & 16) >> 4);
}

__host__ __device__ constexpr float cos32(int q) {
    constexpr float t[16] = {1.0f,
                             0.98078528040323044913f,
                             0.92387953251128675613f,
                             0.83146961230254523708f,
                             0.70710678118654752440f,
                             0.55557023301960222474f,
                             0.38268343236508977173f,
                             0.19509032201612826785f,
                             0.0f,
                             -0.19509032201612826785f,
                             -0.38268343236508977173f,
                             -0.55557023301960222474f,
                             -0.70710678118654752440f,
                             -0.83146961230254523708f,
                             -0.92387953251128675613f,
                             -0.98078528040323044913f};
    return t[q];
}
// sin(2*pi*q/32) = cos(2*pi*(8-q)/32) for q <= 8, cos(2*pi*(q-8)/32) for q in (8,16)
__host__ __device__ constexpr float sin32x(int q) { return q <= 8 ? cos32(8 - q) : cos32(q - 8); }

// ------------------------------------------------------------------------------------------------
// Packed FP32 (sm_100 FFMA2 / FADD2 / FMUL2): a complex value lives in one aligned 64-bit register pair (re, im) and
// every butterfly acts on both halves at once.  ptxas folds the (re, im) swap and the per-half sign of a multiplication
// by +-i into the operand swizzle of the packed instruction (R.F32x2.LO_HI.NP), broadcast constants into its immediate
// slot and broadcast registers into the R.F32 form, so a radix-2 butterfly costs 3 instructions (2 when the twiddle is
// trivial) instead of 6 (4).  Rounding per half is that of the scalar fma/add, so the scalar model in
// tests/fft_model.py describes this code as well.
typedef unsigned long long c64;
__device__ __forceinline__ c64 pk(float a, float b) { c64 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ void upk(c64 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ c64 add2(c64 a, c64 b) { c64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ c64 sub2(c64 a, c64 b) { c64 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ c64 mul2(c64 a, c64 b) { c64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ c64 fma2(c64 a, c64 b, c64 c) {
    c64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ c64 rot_mi(c64 v) { float a, b; upk(v, a, b); return pk(b, -a); }   // -i * v
__device__ __forceinline__ c64 rot_pi(c64 v) { float a, b; upk(v, a, b); return pk(-b, a); }   // +i * v
__device__ __forceinline__ c64 bc2(float c) { return pk(c, c); }                                // (c, c)
__device__ __forceinline__ c64 neg2(c64 v) { float a, b; upk(v, a, b); return pk(-a, -b); }
// Hann phases of a lane's two samples of a row, (cos, cos', sin, sin')(pi (2 lane + e) / 1024), e = 0, 1: the 512-byte table
// process_frame builds the window from (threads 0..31 of the CTA fill it; s_hann is 16-byte aligned)
__device__ __forceinline__ void fill_hann_phases(float2* s_hann, int tid) {
    if (tid < 32) {
        float s0, c0, s1, c1;
        sincospif(static_cast<float>(2 * tid) * (1.0f / 1024.0f), &s0, &c0);
        sincospif(static_cast<float>(2 * tid + 1) * (1.0f / 1024.0f), &s1, &c1);
        reinterpret_cast<float4*>(s_hann)[tid] = make_float4(c0, c1, s0, s1);
    }
}
// cos / sin of r * 2 pi / 32, r = 0..15 (the row phase of the Hann window, see process_frame)
__device__ constexpr float kCos32[16] = {1.0f, 0.9807852804032304f, 0.9238795325112867f, 0.8314696123025452f, 0.7071067811865476f,
                                         0.5555702330196023f, 0.38268343236508984f, 0.19509032201612833f, 0.0f,
                                         -0.1950903220161282f, -0.3826834323650897f, -0.555570233019602f, -0.7071067811865475f,
                                         -0.8314696123025453f, -0.9238795325112867f, -0.9807852804032304f};
__device__ constexpr float kSin32[16] = {0.0f, 0.19509032201612825f, 0.3826834323650898f, 0.5555702330196022f, 0.7071067811865475f,
                                         0.8314696123025452f, 0.9238795325112867f, 0.9807852804032304f, 1.0f,
                                         0.9807852804032304f, 0.9238795325112867f, 0.8314696123025455f, 0.7071067811865476f,
                                         0.5555702330196022f, 0.3826834323650899f, 0.1950903220161286f};
__device__ __forceinline__ c64 conj2(c64 v) { float a, b; upk(v, a, b); return pk(a, -b); }
// v * w for complex v and w = (wr, wi): wr * v + wi * (i v)
__device__ __forceinline__ c64 cmul(c64 v, float wr, float wi) { return fma2(rot_pi(v), bc2(wi), mul2(v, bc2(wr))); }

// radix-2 DIT butterfly (a, b) <- (a + w b, a - w b), w = exp(-2 pi i Q / 32).  Non-trivial twiddles use the "tangent"
// form: w*b = c*(br + t*bi, bi - t*br) with t = s/c when |c| >= |s|, and w*b = s*(t*br + bi, t*bi - br) with t = c/s
// otherwise (|t| <= 1 in both cases), i.e. one packed FMA for the rotation and one each for a + w b and a - w b.
template <int Q>
__device__ __forceinline__ void dit_bfly_p(c64& a, c64& b) {
    if constexpr (Q == 0) {
        const c64 x = sub2(a, b);
        a = add2(a, b); b = x;
    } else if constexpr (Q == 8) {
        const c64 r = rot_mi(b);
        const c64 x = sub2(a, r);
        a = add2(a, r); b = x;
    } else {
        constexpr float c = cos32(Q), sn = sin32x(Q);
        constexpr bool use_c = (c >= 0 ? c : -c) >= sn;
        if constexpr (use_c) {
            constexpr float t = sn / c;
            const c64 p = fma2(rot_mi(b), bc2(t), b);          // (br + t bi, bi - t br)
            b = fma2(p, bc2(-c), a);
            a = fma2(p, bc2(c), a);
        } else {
            constexpr float t = c / sn;
            const c64 p = rot_mi(fma2(rot_pi(b), bc2(t), b));  // (t br + bi, t bi - br)
            b = fma2(p, bc2(-sn), a);
            a = fma2(p, bc2(sn), a);
        }
    }
}
// radix-2 DIT stage of half-size H on the bit-reversed view v[p] = z[brev5(p)]
template <int H>
__device__ __forceinline__ void dit_stage_p(c64 (&z)[32]) {
    sfor<16 / H>([&](auto B) {
        sfor<H>([&](auto J) {
            constexpr int p0 = decltype(B)::value * 2 * H + decltype(J)::value;
            constexpr int i0 = brev5(p0), i1 = brev5(p0 + H);
            constexpr int Q = decltype(J)::value * (16 / H);
            dit_bfly_p<Q>(z[i0], z[i1]);
        });
    });
}
// 32-point complex FFT on packed values: natural-order input x[n] in slot n, output X[k] in slot brev5(k)
// (kFirst = 2 when the caller has already done the first stage, slots r and r + 16)
template <int kFirst = 1>
__device__ __forceinline__ void fft32p(c64 (&z)[32]) {
    if constexpr (kFirst == 1) dit_stage_p<1>(z);
    dit_stage_p<2>(z);
    dit_stage_p<4>(z);
    dit_stage_p<8>(z);
    dit_stage_p<16>(z);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// predicated 16-byte store of a peak record at base[idx] (one mad.wide + one predicated st, no branch)
__device__ __forceinline__ void st_record_if(bool pred, float4* base, unsigned idx, float a, float b, float c, int k) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .u64 ad;\n\t"
        "setp.ne.u32 p, %0, 0;\n\t"
        "mad.wide.u32 ad, %1, 16, %2;\n\t"
        "@p st.global.v4.b32 [ad], {%3, %4, %5, %6};\n\t}"
        :: "r"(static_cast<unsigned>(pred)), "r"(idx), "l"(base), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)),
           "r"(__float_as_uint(c)), "r"(k)
        : "memory");
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ unsigned fkey(float f) {      // order-preserving float -> uint
    const unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(unsigned k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__device__ __forceinline__ int pidx(int k) { return k + 4 * (k >> 5); }   // padded index of bin k in the P tile

// ------------------------------------------------------------------------------------------------
// CTA-wide radix select over the clip's peak magnitudes: key of the element of ascending rank r.
// count_le = number of elements <= that key; has_next / next = the key of rank r + 1 when the select determined it on the
// way (the median of an even count needs both).  All threads must call.  h0, h1, h2 = three 256-int histograms; h0 must be
// all zero and visible to every thread on entry (the caller clears it before its last barrier); `clean` returns one of the
// three, all zero and visible again on return, for the caller's next histogram.
// kor / kand = OR / AND of all keys: only the bits in which the keys differ are examined, 8 per pass from the top, so
// the first pass already spreads over up to 256 bins (keys are order-preserving float bits and a clip's peak
// magnitudes share their leading exponent bits) instead of piling shared-memory atomics onto a handful of bins.
// One barrier per pass: the histograms rotate (the one of pass p + 1 is cleared during pass p: its last readers passed the
// barrier of pass p - 1), and every warp scans the finished histogram itself instead of waiting for warp 0 to publish the
// bucket.  As soon as the surviving bucket holds <= kSelCand keys they are collected (one more scan) and ranked by counting,
// every warp for itself -- typically after one or two of the four passes.
constexpr int kSelCand = 64;
static __device__ unsigned radix_select(const unsigned* keys, int np, int r, int* h0, int* h1, int* h2, int& count_le,
                                        unsigned kor, unsigned kand, unsigned& next, bool& has_next, int*& clean) {
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned diff = kor ^ kand;
    int remaining = 32 - __clz(diff);                       // differing bits are [0, remaining)
    unsigned mask = remaining >= 32 ? 0u : ~((1u << remaining) - 1u);
    unsigned prefix = kand & mask;
    int less = 0, equal = np;
    int* cur = h0; int* nxt = h1; int* third = h2;
    while (remaining > 0 && equal > kSelCand) {
        const int width = min(8, remaining);
        const int shift = remaining - width;
        const unsigned bmask = (1u << width) - 1u;
        for (int i = tid; i < 256; i += kThreads) nxt[i] = 0;
        // four independent loads per thread and step: a long clip's keys live in global memory (np > kKeyCap), and one
        // load in flight per thread made every pass a chain of L2 round trips
        for (int i0 = tid; i0 < np; i0 += 4 * kThreads) {
            unsigned k4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) k4[u] = (i0 + u * kThreads < np) ? keys[i0 + u * kThreads] : 0u;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i0 + u * kThreads < np && (k4[u] & mask) == prefix) atomicAdd(&cur[(k4[u] >> shift) & bmask], 1);
        }
        __syncthreads();
        {
            int loc[8];
            int sum = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) { loc[q] = cur[lane * 8 + q]; sum += loc[q]; }
            int inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            const int exc = inc - sum;
            const unsigned bal = __ballot_sync(0xffffffffu, inc > r);
            const int L = __ffs(bal) - 1;
            int rr = r - exc, cum = 0, sel = 0, eq = 0;
            bool done = false;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (!done) {
                    if (cum + loc[u] > rr) { done = true; sel = u; eq = loc[u]; }
                    else cum += loc[u];
                }
            }
            const int bucket = __shfl_sync(0xffffffffu, lane * 8 + sel, L);
            r = __shfl_sync(0xffffffffu, rr - cum, L);
            less += __shfl_sync(0xffffffffu, exc + cum, L);
            equal = __shfl_sync(0xffffffffu, eq, L);
            prefix |= static_cast<unsigned>(bucket) << shift;
            mask |= bmask << shift;
            remaining = shift;
        }
        int* t = cur; cur = nxt; nxt = third; third = t;
    }
    if (remaining == 0) {                                  // every differing bit consumed: `equal` copies of one key
        clean = cur;                                       // cleared during the last pass (or by the caller), never used
        count_le = less + equal;
        next = prefix;
        has_next = r + 1 < equal;
        return prefix;
    }
    // candidate stage: cur is cleared and unused; cur[0] = count, cur[1 ..] = the bucket's keys in arrival order.  nxt (last
    // read before the previous barrier, if ever) is cleared for the caller meanwhile.
    for (int i = tid; i < 256; i += kThreads) nxt[i] = 0;
    clean = nxt;
    for (int i0 = tid; i0 < np; i0 += 4 * kThreads) {
        unsigned k4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) k4[u] = (i0 + u * kThreads < np) ? keys[i0 + u * kThreads] : 0u;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (i0 + u * kThreads < np && (k4[u] & mask) == prefix) cur[1 + atomicAdd(&cur[0], 1)] = static_cast<int>(k4[u]);
    }
    __syncthreads();
    const int nc = equal;                                   // == cur[0]
    const unsigned* cand = reinterpret_cast<const unsigned*>(cur + 1);
    // lane ranks candidates lane and lane + 32 (ties by list position); the one of rank r is the answer
    const unsigned my0 = lane < nc ? cand[lane] : 0xffffffffu, my1 = lane + 32 < nc ? cand[lane + 32] : 0xffffffffu;
    int rk0 = 0, rk1 = 0;
    for (int j = 0; j < nc; ++j) {
        const unsigned kj = cand[j];
        rk0 += (kj < my0) | ((kj == my0) & (j < lane));
        rk1 += (kj < my1) | ((kj == my1) & (j < lane + 32));
    }
    const bool hit0 = lane < nc && rk0 == r, hit1 = lane + 32 < nc && rk1 == r;
    const bool nx0 = lane < nc && rk0 == r + 1, nx1 = lane + 32 < nc && rk1 == r + 1;
    const unsigned b0 = __ballot_sync(0xffffffffu, hit0), b1 = __ballot_sync(0xffffffffu, hit1);
    const unsigned n0 = __ballot_sync(0xffffffffu, nx0), n1 = __ballot_sync(0xffffffffu, nx1);
    const unsigned key = b0 ? __shfl_sync(0xffffffffu, my0, __ffs(b0) - 1) : __shfl_sync(0xffffffffu, my1, __ffs(b1) - 1);
    has_next = (n0 | n1) != 0u;
    next = n0 ? __shfl_sync(0xffffffffu, my0, __ffs(n0) - 1) : __shfl_sync(0xffffffffu, my1, n1 ? __ffs(n1) - 1 : 0);
    int le = (lane < nc && my0 <= key ? 1 : 0) + (lane + 32 < nc && my1 <= key ? 1 : 0);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) le += __shfl_xor_sync(0xffffffffu, le, o);
    count_le = less + le;
    return key;
}

// D(16x8, f32) += A(16x16, f16, row) * B(16x8, f16, col): warp-level tensor-core MMA, FP32 accumulate
__device__ __forceinline__ void mma_f16(float (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0,
                                        unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ unsigned pack_half2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const unsigned*>(&h);
}
__device__ __forceinline__ void prefetch_l2(const void* ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }
// TMA bulk prefetch of a contiguous global range into L2 (one instruction, no registers, no completion to wait for);
// ptr 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_prefetch_l2(const void* ptr, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ptr), "r"(bytes) : "memory");
}

// Tell L2 that a range of scratch is dead: the 128-byte lines that lie entirely inside [ptr, ptr + bytes) are discarded
// (discard.global.L2), i.e. dropped without a write-back if they are still resident and dirty.  All kThreads threads of the
// CTA call it; the range must not be read again before it is rewritten.
__device__ __forceinline__ void discard_l2_range(const void* ptr, size_t bytes, int tid) {
    const unsigned long long a0 = (reinterpret_cast<unsigned long long>(ptr) + 127ull) & ~127ull;
    const unsigned long long a1 = (reinterpret_cast<unsigned long long>(ptr) + bytes) & ~127ull;
    for (unsigned long long a = a0 + 128ull * tid; a < a1; a += 128ull * kThreads)
        asm volatile("discard.global.L2 [%0], 128;" ::"l"(a) : "memory");
}

// ---- TMA bulk copy (global -> shared) completing on an mbarrier
__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    const long long t0 = clock64();
    for (;;) {
        unsigned done;
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) return;
        if (clock64() - t0 > 4000000000ll) __trap();     // ~2 s: a lost completion must fail loudly, not hang the GPU
    }
}

// ---- tcgen05 (UMMA + tensor memory), used by the SFX_CHROMA_UMMA build of the fused kernel's chroma projection
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s_nx(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {   // no expect_tx
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc(unsigned* smem_dst, unsigned ncols) {      // one warp, converged
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(unsigned taddr, unsigned ncols) {        // one warp, converged
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// shared-memory matrix descriptor of a K-major operand block with the 128-byte swizzle: rows of 128 bytes, 8-row atoms of
// 1 KB (stride byte offset 1024), start address 1 KB aligned plus k * 32 bytes for the k-th 16-element K step of the block
__device__ __forceinline__ unsigned long long umma_desc_sw128(unsigned smem_addr) {
    return static_cast<unsigned long long>((smem_addr & 0x3ffffu) >> 4) | (1ull << 16) | (static_cast<unsigned long long>(1024 >> 4) << 32) |
           (1ull << 46) | (2ull << 61);
}
// D[tmem] (+)= A[smem] . B[smem]^T, FP16 operands, FP32 accumulate; issued by one thread
__device__ __forceinline__ void umma_f16(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {               // arrives on bar when all prior MMAs are done
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 consecutive FP32 columns of this thread's TMEM lane (warp w reads lanes 32*(w%4) .. +31)
__device__ __forceinline__ void tmem_ld32(unsigned taddr, float (&v)[32]) {
    unsigned r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// byte offset, inside a clip's UMMA-image |X|^2 scratch, of the 16-byte chunk c (8 bins) of K block kb of frame t
__device__ __forceinline__ size_t umma_p16_offset(int Tmax, int t, int kb, int c) {
    const int tile = t >> 7, r = t & 127;
    const int na = min(16, ((Tmax + 7) >> 3) - 16 * tile);          // 8-row atoms of this tile (the last tile may be short)
    return static_cast<size_t>(tile) * (16 * 16 * 1024) + static_cast<size_t>(kb) * na * 1024 + (r >> 3) * 1024 + (r & 7) * 128 +
           ((c ^ (r & 7)) << 4);
}

// raw samples of STFT frame t (zero padded) -> re[m1] = x[2m], im[m1] = x[2m+1], m = 32*m1 + lane
__device__ __forceinline__ void load_frame(const float* __restrict__ x, long long n, int t, int lane, bool aligned8,
                                           float (&re)[32], float (&im)[32]) {
    const long long s0 = static_cast<long long>(kHop) * t - kNfft / 2;
    if (s0 >= 0 && s0 + kNfft <= n && aligned8) {
        const float2* src = reinterpret_cast<const float2*>(x + s0) + lane;
#pragma unroll
        for (int m1 = 0; m1 < 32; ++m1) {
            const float2 v = __ldg(src + 32 * m1);
            re[m1] = v.x;
            im[m1] = v.y;
        }
    } else {
#pragma unroll
        for (int m1 = 0; m1 < 32; ++m1) {
            const long long g = s0 + 2 * (32 * m1 + lane);
            re[m1] = (g >= 0 && g < n) ? __ldg(x + g) : 0.0f;
            im[m1] = (g + 1 >= 0 && g + 1 < n) ? __ldg(x + g + 1) : 0.0f;
        }
    }
}

}  // namespace sfx
