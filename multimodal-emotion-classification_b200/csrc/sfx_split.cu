// sfx_split.cu -- the same extractor as sfx_kernels.cu, split into two kernels per chunk of clips:
//   sfx_frames_kernel : phase 1 as a pure stream.  Every warp pulls (clip, frame) items from one global queue; no
//                       CTA-level barrier, no per-clip imbalance.  Per-frame results (FP16 |X|^2 row, log-mel row, peak
//                       records, hop energy, centroid, roll-off, weighted zero crossings, row max) go to the clip's slice.
//   sfx_clips_kernel  : phases 2-3 (tuning estimate, MFCC, tensor-core chroma, pooled row) with one persistent CTA per
//                       clip, its own register budget and 3 CTAs/SM, because these phases are latency-bound.
// Both kernels call the phase functions of sfx_phases.cuh, the same code the fused kernel runs; only where per-frame
// results are kept differs (process_frame<., kSplit = true>).
#include "sfx_phases.cuh"

namespace sfx {

struct Slice {
    __half* gP16; float* gL; float4* gRec; unsigned* gKey; float* gE; float* gNy; float* gInvS;
    float* gCent; float* gRoll; float* gLmax; int* gZc; unsigned char* gBin;
};

__device__ __forceinline__ Slice slice_of(unsigned char* base, int Tmax, int max_pk) {
    Slice s;
    s.gP16 = reinterpret_cast<__half*>(base);
    s.gL = reinterpret_cast<float*>(s.gP16 + static_cast<size_t>(Tmax) * kP16Row);
    s.gRec = reinterpret_cast<float4*>(s.gL + static_cast<size_t>(Tmax) * kMels);
    s.gKey = reinterpret_cast<unsigned*>(s.gRec + static_cast<size_t>(Tmax) * max_pk);
    s.gE = reinterpret_cast<float*>(s.gKey + static_cast<size_t>(Tmax) * max_pk);
    s.gNy = s.gE + Tmax;
    s.gInvS = s.gNy + Tmax;
    s.gCent = s.gInvS + Tmax;
    s.gRoll = s.gCent + Tmax;
    s.gLmax = s.gRoll + Tmax;
    s.gZc = reinterpret_cast<int*>(s.gLmax + Tmax);
    s.gBin = reinterpret_cast<unsigned char*>(s.gZc + Tmax);
    return s;
}

// ---------------------------------------------------------------------------------------------- prep
// zero the queues / peak counters of the chunk and, for ragged batches, build the exclusive prefix of frame counts
__global__ void __launch_bounds__(1024) sfx_prep_kernel(const SplitParams q) {
    __shared__ int s_w[32];
    const int c = threadIdx.x, lane = c & 31, warp = c >> 5;
    int* hdr = reinterpret_cast<int*>(q.p.ws);
    int* npk = reinterpret_cast<int*>(q.p.ws + kSplitNpkOff);
    int* off = reinterpret_cast<int*>(q.p.ws + kSplitOffOff);
    if (c < q.nclips) npk[c] = 0;
    int T = 0;
    if (c < q.nclips) {
        const long long n = clip_samples(q.p, q.chunk0 + c);
        T = n > 0 ? 1 + static_cast<int>(n / kHop) : 0;
    }
    int inc = T;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = s_w[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += v;
        }
        s_w[lane] = w;
    }
    __syncthreads();
    const int base = warp ? s_w[warp - 1] : 0;
    if (c < q.nclips) off[c] = base + inc - T;
    if (c == q.nclips - 1) { off[q.nclips] = base + inc; hdr[2] = base + inc; }
    if (c == 0) { hdr[0] = 0; hdr[1] = 0; }
}

// ---------------------------------------------------------------------------------------------- frames
template <bool kDebug>
__global__ void __launch_bounds__(kThreads, 2) sfx_frames_kernel(const SplitParams q) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* s_hann = reinterpret_cast<float2*>(smem_raw);
    float2* s_tw1 = s_hann + 64;                     // (s_hann: 32 lanes x (cos, cos', sin, sin'), fill_hann_phases)
    float2* s_tw2 = s_tw1 + 1024;
    float2* s_melab = s_tw2 + 512;
    float* s_ex = reinterpret_cast<float*>(s_melab + 17 * 64);               // [kWarps][kExFloats]

    const Params& p = q.p;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const DevTables& tb = p.tb;
    fill_hann_phases(s_hann, tid);
    for (int i = tid; i < 1024; i += kThreads) {
        const int r = i >> 5, l = i & 31;
        const int d = (r >> 1) * 64 + 2 * l + (r & 1);
        s_tw1[d] = tb.tw1[i];
        if (i < 512) s_tw2[d] = tb.tw2[i];
    }
    for (int i = tid; i < 33 * 32; i += kThreads) {
        const int r = i >> 5, l = i & 31;
        s_melab[(r >> 1) * 64 + 2 * l + (r & 1)] = tb.mel_ab[i];
    }
    const unsigned mel_mask = tb.mel_mask[lane];
    const int mel_ps = tb.mel_ps;
    int msrc[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const int* m3 = tb.mel_src + (32 * s + lane) * 3;
        msrc[s] = m3[0] | (m3[1] << 10) | (m3[2] << 20);
    }
    __syncthreads();

    int* queue = reinterpret_cast<int*>(p.ws);
    int* npk_all = reinterpret_cast<int*>(p.ws + kSplitNpkOff);
    const int* off = reinterpret_cast<const int*>(p.ws + kSplitOffOff);
    const int total = queue[2];
    FrameSmem fs;
    fs.s_hann = s_hann; fs.s_tw1 = s_tw1; fs.s_tw2 = s_tw2; fs.s_melab = s_melab;
    fs.Pb = s_ex + warp * kExFloats;
    fs.ex = reinterpret_cast<float2*>(fs.Pb);
    fs.part = fs.Pb + kPartOff;
    fs.mel_mask = mel_mask; fs.mel_ps = mel_ps;
#pragma unroll
    for (int s = 0; s < 4; ++s) fs.msrc[s] = msrc[s];
    fs.s_msrc = nullptr;
    fs.bin_hz = static_cast<float>(static_cast<double>(tb.sr) / kNfft);
    fs.aligned8 = p.aligned8 != 0;

    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(queue, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= total) break;
        int c;
        if (q.T_uniform > 0) {
            c = item / q.T_uniform;
        } else {                                   // largest c with off[c] <= item
            int lo = 0, hi = q.nclips;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (off[mid] <= item) lo = mid; else hi = mid;
            }
            c = lo;
        }
        const int t = item - (q.T_uniform > 0 ? c * q.T_uniform : off[c]);
        const int clip = q.chunk0 + c;
        const long long n = clip_samples(p, clip);
        const int T = 1 + static_cast<int>(n / kHop);
        const float* x = p.wave + static_cast<long long>(clip) * p.row_stride;
        const Slice sl = slice_of(p.ws + kSplitHeader + static_cast<size_t>(c) * p.cta_scratch_bytes, p.Tmax, p.max_pk);
        FrameOut fo;
        fo.gP16 = sl.gP16; fo.gL = sl.gL; fo.gRec = sl.gRec; fo.gE = sl.gE; fo.gNy = sl.gNy; fo.gInvS = sl.gInvS;
        fo.npk = npk_all + c; fo.gSeg = nullptr; fo.s_wacc = nullptr; fo.s_f = nullptr;
        fo.gCent = sl.gCent; fo.gRoll = sl.gRoll; fo.gLmax = sl.gLmax; fo.gZc = sl.gZc; fo.gFv = nullptr; fo.cursor = nullptr; fo.s_lm = nullptr; fo.s_lmin = nullptr;
        int unused_zc = 0, unused_cnt = 0;
        process_frame<kDebug, kModeSplit>(p, tb, fs, fo, x, n, T, t, clip, lane, warp, unused_zc, unused_cnt);
    }
}

// ---------------------------------------------------------------------------------------------- clips
template <bool kDebug>
__global__ void __launch_bounds__(kThreads, 2) sfx_clips_kernel(const SplitParams q) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_ex = reinterpret_cast<float*>(smem_raw);                        // [kWarps][kExFloats]
    double* s_pool = reinterpret_cast<double*>(s_ex + kWarps * kExFloats);   // [256]
    double* s_wacc = s_pool + 256;                                           // [kWarps][16]
    double* s_edges = s_wacc + kWarps * 16;                                  // [104]
    unsigned long long* s_mbar = reinterpret_cast<unsigned long long*>(s_edges + 104);
    int* s_hist = reinterpret_cast<int*>(s_mbar + 1);                        // [256]
    int* s_i = s_hist + 256;                                                 // [32]
    float* s_f = reinterpret_cast<float*>(s_i + 32);                         // [32]

    const Params& p = q.p;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const DevTables& tb = p.tb;
    for (int i = tid; i <= kTunings; i += kThreads) s_edges[i] = tb.edges[i];
    if (tid == 0) mbar_init(s_mbar, 1);
    unsigned bank_parity = 0;
    int* queue = reinterpret_cast<int*>(p.ws);
    const int* npk_all = reinterpret_cast<const int*>(p.ws + kSplitNpkOff);
    const ClipSmem cs{s_ex, s_pool, s_wacc, s_edges, s_mbar, s_hist, s_i, s_f, nullptr, nullptr, nullptr, 0u};

    for (;;) {
        __syncthreads();
        if (tid == 0) s_i[0] = atomicAdd(queue + 1, 1);
        __syncthreads();
        const int c = s_i[0];
        if (c >= q.nclips) break;
        const int clip = q.chunk0 + c;
        const long long n = clip_samples(p, clip);
        float* out = p.out + static_cast<long long>(clip) * p.out_stride;
        if (n <= 0) {
            for (int i = tid; i < p.n_mfcc + 16; i += kThreads) out[i] = __int_as_float(0x7fc00000);
            continue;
        }
        const int T = 1 + static_cast<int>(n / kHop);
        const Slice sl = slice_of(p.ws + kSplitHeader + static_cast<size_t>(c) * p.cta_scratch_bytes, p.Tmax, p.max_pk);
        // per-clip sums of the per-frame descriptors written by the frames kernel (fixed thread -> frame assignment:
        // deterministic) into the slots the phases below read
        {
            double sc = 0.0, sr = 0.0;
            int sz = 0;
            float lm = -FLT_MAX, lastnz = -1.0f;
            for (int t = tid; t < T; t += kThreads) {
                sc += static_cast<double>(sl.gCent[t]);
                sr += static_cast<double>(sl.gRoll[t]);
                sz += sl.gZc[t];
                lm = fmaxf(lm, sl.gLmax[t]);
                if (sl.gInvS[t] >= 0.0f) lastnz = static_cast<float>(t);     // the frames kernel marks all-zero frames with -1
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                sc += __shfl_xor_sync(0xffffffffu, sc, o);
                sr += __shfl_xor_sync(0xffffffffu, sr, o);
            }
            sz = warp_sum_i(sz);
            lm = warp_max(lm);
            lastnz = warp_max(lastnz);
            if (lane == 0) { s_wacc[warp * 16 + 0] = sc; s_wacc[warp * 16 + 1] = sr; s_i[8 + warp] = sz; s_f[warp] = lm;
                             s_f[8 + warp] = lastnz; }       // last non-zero frame: the chroma projection stops there (as in the fused kernel)
            if (tid < kWarps) s_i[20 + tid] = tid == 0 ? npk_all[c] : 0;     // all records in segment 0
            if (tid == 0) { s_i[17] = 0; s_i[18] = 0; s_i[19] = -1; }
        }
        __syncthreads();

        const ClipSlice cl{sl.gP16, sl.gL, sl.gRec, sl.gKey, sl.gE, sl.gNy, sl.gInvS, sl.gBin, 0};
        clip_tail<kDebug>(p, tb, cs, cl, clip, T, out, bank_parity, tid, lane, warp);
    }
}

// ---------------------------------------------------------------------------------------------- host
size_t smem_frames() { return sizeof(float2) * (64 + 1536 + 17 * 64) + sizeof(float) * kWarps * kExFloats; }
size_t smem_clips() {
    return sizeof(float) * kWarps * kExFloats + sizeof(double) * (256 + kWarps * 16 + 104 + 1) + sizeof(int) * (256 + 32) +
           sizeof(float) * 32;
}

cudaError_t configure_split(int* frames_per_sm, int* clips_per_sm) {
    cudaError_t e;
    const int sf = static_cast<int>(smem_frames()), sc = static_cast<int>(smem_clips());
    if ((e = cudaFuncSetAttribute(sfx_frames_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sf)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(sfx_frames_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sf)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(sfx_clips_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sc)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(sfx_clips_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sc)) != cudaSuccess) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(frames_per_sm, sfx_frames_kernel<false>, kThreads, sf)) != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(clips_per_sm, sfx_clips_kernel<false>, kThreads, sc);
}

// one chunk: prep -> frames -> clips on `stream`
cudaError_t launch_split_chunk(const SplitParams& q, int grid_frames, int grid_clips, bool debug, cudaStream_t stream) {
    sfx_prep_kernel<<<1, 1024, 0, stream>>>(q);
    if (debug) sfx_frames_kernel<true><<<grid_frames, kThreads, smem_frames(), stream>>>(q);
    else       sfx_frames_kernel<false><<<grid_frames, kThreads, smem_frames(), stream>>>(q);
    if (debug) sfx_clips_kernel<true><<<grid_clips, kThreads, smem_clips(), stream>>>(q);
    else       sfx_clips_kernel<false><<<grid_clips, kThreads, smem_clips(), stream>>>(q);
    return cudaGetLastError();
}

}  // namespace sfx
