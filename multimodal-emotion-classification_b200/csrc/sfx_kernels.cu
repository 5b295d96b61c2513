// sfx_kernels.cu -- sm_100a kernels of the batched speech feature extractor.
//
// One persistent CTA per clip (dynamic clip queue), 8 warps, 2 CTAs per SM.  Replaces, for a whole batch,
// the per-clip librosa calls of the reference's preprocessing/audio_preprocessing.py:22-37:
//   phase 1 (warp per STFT frame): framing + Hann + 2048-pt real FFT (1024-pt complex FFT as two register-resident
//           radix-32 passes on packed FP32 -- FFMA2 / FADD2, a complex value per register pair -- with one
//           shared-memory transpose) -> |X|^2; rms / zero crossings from the raw samples; |X| centroid + 0.85
//           roll-off (warp scan); piptrack peaks (librosa.piptrack on the POWER spectrum) ballot-compacted into the
//           warp's own segment of the clip's record buffer; sparse Slaney mel projection + 10*log10.  FP16-scaled
//           |X|^2 rows and log-mel rows go to the CTA's scratch slice.
//   phase 2 (CTA): estimate_tuning = median of peak magnitudes (radix select over the differing key bits) ->
//           100-bin histogram arg-max, per-peak arithmetic in guarded fast forms that reproduce librosa's float64 /
//           float32 dtype trail exactly; global log-mel max for power_to_db(top_db=80).
//   phase 3 (CTA): clamp + frame-mean of log-mel, DCT-II (float64) -> MFCC; chroma projection with the tuning's
//           filter bank (TMA-staged, tensor-core MMA), per-frame inf-norm, frame mean; pooled spectral descriptors.
// The phase code itself lives in sfx_phases.cuh (shared with the frame-parallel pipeline of sfx_split.cu).
// Output row: [mfcc_0..n_mfcc-1 | chroma C..B | zcr, centroid, rolloff, rms]  (reference :45-46).
#include "sfx_phases.cuh"

namespace sfx {

// ------------------------------------------------------------------------------------------------
// kUmma: the chroma projection of phase 3b on tcgen05 (UMMA, accumulator in tensor memory, operands as bulk-copied
// shared-memory images) instead of mma.sync: pipeline mode 4, an A/B of the two tensor paths in one library.
template <bool kDebug, bool kUmma>
__global__ void __launch_bounds__(kThreads, 2) sfx_extract_kernel(const Params p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* s_hann = reinterpret_cast<float2*>(smem_raw);
    float2* s_tw1 = s_hann + 64;                                             // (s_hann: 32 lanes x (cos, cos', sin, sin'))
    float2* s_tw2 = s_tw1 + 1024;
    float2* s_melab = s_tw2 + 512;                                           // [33*32]  (tw2: rows k2 < 16 only)
    float* s_ex = reinterpret_cast<float*>(s_melab + 17 * 64);               // [kWarps][kExFloats]
    double* s_pool = reinterpret_cast<double*>(s_ex + kWarps * kExFloats);   // [256]
    double* s_wacc = s_pool + 256;                                           // [kWarps][16]
    double* s_edges = s_wacc + kWarps * 16;                                  // [104]
    unsigned long long* s_mbar = reinterpret_cast<unsigned long long*>(s_edges + 104);   // chroma-bank TMA barrier
    int* s_hist = reinterpret_cast<int*>(s_mbar + 1);                        // [256]
    int* s_i = s_hist + 256;                                                 // [32]
    float* s_f = reinterpret_cast<float*>(s_i + 32);                         // [32]
    float* s_lmin = s_f + 32;                                                // [kWarps][32]
    double* s_lm = reinterpret_cast<double*>(s_lmin + kWarps * 32);          // [kWarps][128]
    unsigned long long* s_ubar = reinterpret_cast<unsigned long long*>(s_lm + kWarps * kMels);   // [8] UMMA build: mbarriers
    unsigned* s_tmem = reinterpret_cast<unsigned*>(s_ubar + 7);              // tensor-memory base address
    int* s_msrc = reinterpret_cast<int*>(s_ubar + 8);                        // [128] partial-sum slot words of the mel filters

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const DevTables& tb = p.tb;

    // tables are stored row-paired: element (row r, lane) sits at float2 index (r/2)*64 + 2*lane + (r&1), so one LDS.128
    // serves rows r, r+1 of a lane (conflict-free: consecutive lanes are 16 B apart)
    for (int i = tid; i < 1024; i += kThreads) {
        const int r = i >> 5, l = i & 31;
        const int d = (r >> 1) * 64 + 2 * l + (r & 1);
        s_tw1[d] = tb.tw1[i];
        if (i < 512) s_tw2[d] = tb.tw2[i];
    }
    for (int i = tid; i < 33 * 32; i += kThreads) {
        const int r = i >> 5, l = i & 31;
        s_melab[(r >> 1) * 64 + 2 * l + (r & 1)] = tb.mel_ab[i];      // row 32 (Nyquist) lands at 16*64 + 2*lane
    }
    fill_hann_phases(s_hann, tid);
    for (int i = tid; i <= kTunings; i += kThreads) s_edges[i] = tb.edges[i];
    if (tid == 0) mbar_init(s_mbar, 1);
    UmmaState us;
    if constexpr (kUmma) {
        if (tid == 0) {
            for (int i = 0; i < 7; ++i) mbar_init(s_ubar + i, 1);
        }
        if (warp == 0) tmem_alloc(s_tmem, 32);             // 32 FP32 columns x 128 lanes: the chroma accumulator tile
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    unsigned bank_parity = 0;
    const unsigned mel_mask = tb.mel_mask[lane];
    const int mel_ps = tb.mel_ps;
    // 3 x 10-bit partial-sum slots per mel filter, read from shared memory once per frame (kept in registers they were
    // spilled: the frame loop has none to spare)
    for (int m = tid; m < kMels; m += kThreads) {
        const int* q = tb.mel_src + m * 3;
        s_msrc[m] = q[0] | (q[1] << 10) | (q[2] << 20);
    }

    // scratch of this CTA: its rows inside the two blocks that hold every CTA's FP16 |X|^2 / log-mel rows (the L2-pinned
    // part of the workspace), and its slice with the rest
    unsigned char* rows0 = p.ws + kWsHeader;
    __half* gP16 = reinterpret_cast<__half*>(rows0 + static_cast<size_t>(blockIdx.x) * p.cta_p16_bytes);           // [Tmax][kP16Stride]
    float* gL = reinterpret_cast<float*>(rows0 + static_cast<size_t>(gridDim.x) * p.cta_p16_bytes +
                                         static_cast<size_t>(blockIdx.x) * p.cta_lm_bytes);                        // [Tmax][128]
    unsigned char* slice = rows0 + static_cast<size_t>(gridDim.x) * (p.cta_p16_bytes + p.cta_lm_bytes) +
                           static_cast<size_t>(blockIdx.x) * p.cta_scratch_bytes;
    float4* gRec = reinterpret_cast<float4*>(slice);
    const int seg_cap = rec_frames(p.Tmax) / kWarps * p.max_pk;               // peak records per warp segment
    unsigned* gKey = reinterpret_cast<unsigned*>(gRec + static_cast<size_t>(kWarps) * seg_cap);
    float* gE = reinterpret_cast<float*>(gKey + static_cast<size_t>(p.Tmax) * p.max_pk);        // hop energies [Tmax]
    float* gNy = gE + p.Tmax;                                                                    // scaled Nyquist |X|^2 [Tmax]
    float* gInvS = gNy + p.Tmax;                                                                 // 1 / row scale [Tmax]
    unsigned char* gBin = reinterpret_cast<unsigned char*>(gInvS + p.Tmax);
    int* counter = reinterpret_cast<int*>(p.ws);

    FrameSmem fs;
    fs.s_hann = s_hann; fs.s_tw1 = s_tw1; fs.s_tw2 = s_tw2; fs.s_melab = s_melab;
    fs.Pb = s_ex + warp * kExFloats;                   // this warp's exchange / |X|^2 tile
    fs.ex = reinterpret_cast<float2*>(fs.Pb);
    fs.part = fs.Pb + kPartOff;                        // mel partial sums [32][mel_ps] + zero slot
    fs.mel_mask = mel_mask; fs.mel_ps = mel_ps;
    fs.s_msrc = s_msrc;
    fs.bin_hz = static_cast<float>(static_cast<double>(tb.sr) / kNfft);
    fs.aligned8 = p.aligned8 != 0;
    FrameOut fo;
    fo.gP16 = gP16; fo.gL = gL; fo.gRec = gRec; fo.gE = gE; fo.gNy = gNy; fo.gInvS = gInvS;
    fo.npk = nullptr; fo.gSeg = gRec + static_cast<size_t>(warp) * seg_cap; fo.s_wacc = s_wacc; fo.s_f = s_f;
    fo.s_lm = s_lm + warp * kMels; fo.s_lmin = s_lmin + warp * 32;
    fo.gCent = nullptr; fo.gRoll = nullptr; fo.gLmax = nullptr; fo.gZc = nullptr; fo.gFv = nullptr; fo.cursor = nullptr;
    const ClipSmem cs{s_ex, s_pool, s_wacc, s_edges, s_mbar, s_hist, s_i, s_f, s_lm, s_lmin, s_ubar, kUmma ? *s_tmem : 0u};
    const ClipSlice sl{gP16, gL, gRec, gKey, gE, gNy, gInvS, gBin, seg_cap};

    // The clip queue: s_i[1] = queue index of the CTA's next clip.  Thread 0 draws it while the tail of the current clip runs
    // (the global atomic's round trip used to sit between two barriers at the top of the loop, with 255 threads waiting) and
    // publishes it behind the tail; one barrier per clip separates the clips.
    if (tid == 0) s_i[1] = atomicAdd(counter, 1);
    for (;;) {
        __syncthreads();                           // the previous clip's tail is done with the shared state; s_i[1] is visible
        const int qi = s_i[1];
        if (qi >= p.B) break;
        const int clip = p.order ? p.order[qi] : qi;
        const long long n = clip_samples(p, clip);
        float* out = p.out + static_cast<long long>(clip) * p.out_stride;
        if (n <= 0) {
            for (int i = tid; i < p.n_mfcc + 16; i += kThreads) out[i] = __int_as_float(0x7fc00000);
            __syncthreads();                       // every thread has read s_i[1]
            if (tid == 0) s_i[1] = atomicAdd(counter, 1);
            continue;
        }
        if (tid == 0) { s_i[17] = 0; s_i[18] = 0; s_i[19] = -1; }      // (published by the barrier behind the frames)
        const int T = 1 + static_cast<int>(n / kHop);
        const float* x = p.wave + static_cast<long long>(clip) * p.row_stride;

        // ===================================== phase 1: frames =====================================
        // per-warp running sums live in shared memory (s_wacc[warp][0..2] = centroid, rolloff, rms; s_f = log-mel max)
        if (lane < 2) s_wacc[warp * 16 + lane] = 0.0;
        if (lane == 0) { s_f[warp] = -FLT_MAX; s_f[8 + warp] = -1.0f; s_f[16 + warp] = 0.0f; }
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) fo.s_lm[32 * s4 + lane] = 0.0;
        fo.s_lmin[lane] = FLT_MAX;
        int acc_zc = 0, wcount = 0;
        __syncwarp();

#ifdef SFX_FUSED_DIAG
        const long long fprof_f0 = clock64();
#endif
        for (int t = warp; t < T; t += kWarps)
            process_frame<kDebug, kUmma ? kModeFusedUmma : kModeFused>(p, tb, fs, fo, x, n, T, t, clip, lane, warp, acc_zc, wcount);

        // per-warp partials
        {
            const int zc = warp_sum_i(acc_zc);
            if (lane == 0) { s_i[8 + warp] = zc; s_i[20 + warp] = wcount; }
        }
        __syncthreads();

#ifdef SFX_FUSED_DIAG
        if (tid == 0) {
            atomicAdd(&g_fprof[0], static_cast<unsigned long long>(clock64() - fprof_f0));
            atomicAdd(&g_fprof[7], 1ull);
        }
        const long long fprof_t0 = clock64();
#endif
        int next_q = 0;
        if (tid == 0) next_q = atomicAdd(counter, 1);                  // every thread has read s_i[1]; stored behind the tail
        clip_tail<kDebug, kUmma, true>(p, tb, cs, sl, clip, T, out, bank_parity, tid, lane, warp, &us);
        if (tid == 0) s_i[1] = next_q;
#ifdef SFX_FUSED_DIAG
        if (tid == 0) atomicAdd(&g_fprof[6], static_cast<unsigned long long>(clock64() - fprof_t0));
#endif
    }
    if constexpr (kUmma) {
        tc_fence_before();
        __syncthreads();
        if (warp == 0) tmem_dealloc(cs.tmem, 32);
    }
}

#ifdef SFX_FUSED_DIAG
extern "C" int sfx_fused_prof(unsigned long long* out, int reset) {
    int rc = static_cast<int>(cudaMemcpyFromSymbol(out, g_fprof, sizeof(g_fprof)));
    if (reset) {
        unsigned long long z[16] = {};
        rc |= static_cast<int>(cudaMemcpyToSymbol(g_fprof, z, sizeof(z)));
    }
    return rc;
}
#endif

// ------------------------------------------------------------------------------------------------
// Longest-processing-time-first order of a ragged batch: one CTA, counting sort of the clips by a 256-level logarithmic
// key of their frame count (8 levels per octave), longest first.  With one persistent CTA per clip, a 60 s clip that is
// pulled from the queue last would otherwise run alone at the end of the launch.
__global__ void __launch_bounds__(1024) sfx_order_kernel(const int32_t* __restrict__ lengths, const int B,
                                                         int* __restrict__ order) {
    __shared__ int s_cnt[256], s_pos[256];
    const int tid = threadIdx.x;
    auto bucket = [](int n) {
        const unsigned T = n > 0 ? 1u + static_cast<unsigned>(n) / kHop : 0u;
        unsigned key = T;
        if (T >= 8u) {
            const int e = 31 - __clz(T);
            key = (static_cast<unsigned>(e) << 3) | ((T >> (e - 3)) & 7u);
        }
        return 255 - static_cast<int>(min(key, 255u));
    };
    if (tid < 256) s_cnt[tid] = 0;
    __syncthreads();
    for (int i = tid; i < B; i += 1024) atomicAdd(&s_cnt[bucket(lengths[i])], 1);
    __syncthreads();
    if (tid < 32) {                                  // exclusive scan of the 256 counters by one warp
        int loc[8], sum = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) { loc[q] = s_cnt[tid * 8 + q]; sum += loc[q]; }
        int inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (tid >= o) inc += v;
        }
        int run = inc - sum;
#pragma unroll
        for (int q = 0; q < 8; ++q) { s_pos[tid * 8 + q] = run; run += loc[q]; }
    }
    __syncthreads();
    for (int i = tid; i < B; i += 1024) order[atomicAdd(&s_pos[bucket(lengths[i])], 1)] = i;
}

cudaError_t launch_order(const int32_t* lengths, int B, int* order, cudaStream_t stream) {
    sfx_order_kernel<<<1, 1024, 0, stream>>>(lengths, B, order);
    return cudaGetLastError();
}

size_t smem_bytes() {
    return sizeof(float2) * (64 + 1536 + 17 * 64) + sizeof(float) * kWarps * kExFloats +
           sizeof(double) * (256 + kWarps * 16 + 104 + 1) + sizeof(int) * (256 + 32) + sizeof(float) * (32 + kWarps * 32) +
           sizeof(double) * kWarps * kMels + sizeof(unsigned long long) * 8 + sizeof(int) * kMels;
}

cudaError_t configure_kernels(int* blocks_per_sm) {
    const int smem = static_cast<int>(smem_bytes());
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(sfx_extract_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(sfx_extract_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(sfx_extract_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(sfx_extract_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
    // (the occupancy calculator answers 1 for the tcgen05 instantiations although two of their CTAs do share an SM -- 2 x 32
    //  of the 512 tensor-memory columns, same registers and shared memory; the mma.sync instantiation's answer is used for both)
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, sfx_extract_kernel<false, false>, kThreads, smem);
}

cudaError_t launch_extract(const Params& p, int grid, bool debug, bool umma, cudaStream_t stream) {
    const size_t smem = smem_bytes();
    if (umma) {
        if (debug) sfx_extract_kernel<true, true><<<grid, kThreads, smem, stream>>>(p);
        else       sfx_extract_kernel<false, true><<<grid, kThreads, smem, stream>>>(p);
    } else {
        if (debug) sfx_extract_kernel<true, false><<<grid, kThreads, smem, stream>>>(p);
        else       sfx_extract_kernel<false, false><<<grid, kThreads, smem, stream>>>(p);
    }
    return cudaGetLastError();
}

}  // namespace sfx
