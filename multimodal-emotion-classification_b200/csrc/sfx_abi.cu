// sfx_abi.cu -- extern "C" boundary of libsfx_b200.so (declared in include/sfx.h).
// Host-only logic: per-device table upload, workspace sizing, launch, and the host-buffer pipeline.
#include <cuda_runtime.h>
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "sfx_internal.h"

namespace {

constexpr int kMaxDev = 16;
constexpr int kHostStreams = 3;

struct HostPath {               // cached buffers of sfx_extract_host
    cudaStream_t stream[kHostStreams] = {};
    float* d_wave[kHostStreams] = {};
    int16_t* d_pcm[kHostStreams] = {};   // raw PCM16 rows of sfx_extract_host_pcm16
    int* d_len[kHostStreams] = {};
    float* d_out[kHostStreams] = {};
    void* d_ws[kHostStreams] = {};
    float* h_stage[kHostStreams] = {};   // pinned staging for pageable input
    float* h_out[kHostStreams] = {};     // pinned staging for pageable output
    cudaEvent_t ev_done[kHostStreams] = {};
    size_t wave_elems = 0, ws_bytes = 0, out_elems = 0;
    int chunk = 0;
};

struct TableSet {               // one per sample rate
    int sr = 0;
    sfx::DevTables tb{};
};

struct DevCtx {
    bool ready = false;
    int sm_count = 0, blocks_per_sm = 0, grid_max = 0;
    int frames_per_sm = 0, clips_per_sm = 0;      // split pipeline occupancies
    int stream_per_sm = 0;                        // stream pipeline: CTAs per SM (1: a 16-warp CTA owns the register file)
    int max_pk = 0;             // max over table sets of local maxima that fit the piptrack bin range (multiple of 4)
    size_t l2_persist = 0;      // bytes of L2 set aside for persisting accesses (0: none), l2_window = largest policy window
    size_t l2_window = 0;
    std::vector<TableSet> sets;
    std::vector<void*> allocs;
    HostPath hp;
};

const TableSet* find_set(const DevCtx& c, int sr) {
    for (const TableSet& t : c.sets)
        if (t.sr == sr) return &t;
    return nullptr;
}

DevCtx g_ctx[kMaxDev];
std::mutex g_tab_mu;            // guards DevCtx::{ready, sets, occupancies, max_pk, allocs}: short critical sections only
std::mutex g_host_mu;           // guards DevCtx::hp (the cached buffers of the host-buffer pipeline) for a whole call
thread_local std::string g_err;
thread_local int g_last_launches = 1;

// What a launch needs from the device context, copied under g_tab_mu (a concurrent sfx_init_tables for another sample
// rate may reallocate DevCtx::sets, so no pointer into it is kept).
struct LaunchCtx {
    sfx::DevTables tb{};
    int sm_count = 0, grid_max = 0, frames_per_sm = 0, clips_per_sm = 0, stream_per_sm = 0, max_pk = 0;
    size_t l2_persist = 0, l2_window = 0;
};
bool snapshot(int device, int sr, LaunchCtx* out) {
    std::lock_guard<std::mutex> lk(g_tab_mu);
    const DevCtx& c = g_ctx[device];
    const TableSet* ts = c.ready ? find_set(c, sr) : nullptr;
    if (!ts) return false;
    out->tb = ts->tb;
    out->sm_count = c.sm_count; out->grid_max = c.grid_max; out->frames_per_sm = c.frames_per_sm;
    out->clips_per_sm = c.clips_per_sm; out->stream_per_sm = c.stream_per_sm; out->max_pk = c.max_pk;
    out->l2_persist = c.l2_persist; out->l2_window = c.l2_window;
    return true;
}

// Pipeline choice: 0 = auto, 1 = fused, 2 = split, 3 = stream.
// Initial value from the environment variable SFX_PIPELINE (auto | fused | split | stream); sfx_set_pipeline() overrides it.
// Auto mode uses the frame-parallel split pipeline when the whole batch fits one chunk and has at most kAutoSplitMaxB clips:
// measured on 3 s clips (tools/batch_sweep.py) it is 4-15 % faster than the persistent kernels up to 1 024 clips and slower
// from 1 440 on.  Ragged batches switch at 256: a few very long clips serialise its per-chunk kernels.  Larger batches take
// kAutoLarge.
constexpr int kAutoSplitMaxB = 1024;
constexpr int kAutoSplitMaxBRagged = 256;
constexpr int kPipeFused = 1, kPipeSplit = 2, kPipeStream = 3, kPipeFusedUmma = 4;
#ifndef SFX_AUTO_LARGE
#define SFX_AUTO_LARGE 1        // measured (tools/ab_modes.py, bench mix): fused 1.78 M clips/s, stream 1.42 M
#endif
constexpr int kAutoLarge = SFX_AUTO_LARGE;
std::atomic<int> g_pipeline{[] {
    const char* e = std::getenv("SFX_PIPELINE");
    if (e && std::strcmp(e, "split") == 0) return kPipeSplit;
    if (e && std::strcmp(e, "fused") == 0) return kPipeFused;
    if (e && std::strcmp(e, "stream") == 0) return kPipeStream;
    if (e && std::strcmp(e, "fused_umma") == 0) return kPipeFusedUmma;
    return 0;
}()};
// L2 pinning of the fused kernel's read-back rows: SFX_L2_PIN = 0 (off) | 1 (FP16 |X|^2 rows, default) | 2 (+ log-mel rows)
const int g_l2_mode = [] {
    const char* e = std::getenv("SFX_L2_PIN");
    return e ? std::atoi(e) : 1;
}();
int choose_pipeline(int mode, int B, bool ragged, size_t split_cap) {
    if (mode != 0) return mode;
    if (B <= (ragged ? kAutoSplitMaxBRagged : kAutoSplitMaxB) && static_cast<size_t>(B) <= split_cap) return kPipeSplit;
    return kAutoLarge;
}

// clips per chunk of the split pipeline: at most kSplitChunkMax, at most ~1 GiB of slices
int split_chunk_for(size_t slice) {
    const size_t budget = size_t(1) << 30;
    size_t c = budget / slice;
    if (c < 1) c = 1;
    if (c > static_cast<size_t>(sfx::kSplitChunkMax)) c = sfx::kSplitChunkMax;
    return static_cast<int>(c);
}

int fail(int code, const std::string& msg) { g_err = msg; return code; }
int cuda_fail(cudaError_t e, const char* what) {
    return fail(SFX_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CK(call)                                              \
    do {                                                      \
        cudaError_t e__ = (call);                             \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)

template <class T>
int upload(DevCtx& c, const T* host, size_t count, const T** dev) {
    void* d = nullptr;
    CK(cudaMalloc(&d, count * sizeof(T)));
    c.allocs.push_back(d);
    CK(cudaMemcpy(d, host, count * sizeof(T), cudaMemcpyHostToDevice));
    *dev = static_cast<const T*>(d);
    return SFX_OK;
}

void free_host_path(HostPath& hp) {
    for (int s = 0; s < kHostStreams; ++s) {
        if (hp.d_wave[s]) cudaFree(hp.d_wave[s]);
        if (hp.d_pcm[s]) cudaFree(hp.d_pcm[s]);
        if (hp.d_len[s]) cudaFree(hp.d_len[s]);
        if (hp.d_out[s]) cudaFree(hp.d_out[s]);
        if (hp.d_ws[s]) cudaFree(hp.d_ws[s]);
        if (hp.h_stage[s]) cudaFreeHost(hp.h_stage[s]);
        if (hp.h_out[s]) cudaFreeHost(hp.h_out[s]);
        if (hp.ev_done[s]) cudaEventDestroy(hp.ev_done[s]);
        if (hp.stream[s]) cudaStreamDestroy(hp.stream[s]);
    }
    hp = HostPath{};
}

bool is_pinned(const void* p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

// PCM16 -> float32 exactly as libsndfile / soundfile do it for librosa.load: x / 32768 (two samples per thread)
__global__ void pcm16_to_f32_kernel(const int16_t* __restrict__ src, float* __restrict__ dst, long long pairs) {
    const long long step = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < pairs; i += step) {
        const short2 v = reinterpret_cast<const short2*>(src)[i];
        reinterpret_cast<float2*>(dst)[i] = make_float2(static_cast<float>(v.x) * (1.0f / 32768.0f),
                                                        static_cast<float>(v.y) * (1.0f / 32768.0f));
    }
}

// bytes of workspace `pipe` needs for B clips of at most max_samples samples (B <= 0: any batch size)
size_t pipeline_ws_bytes(const LaunchCtx& c, int pipe, int64_t max_samples, int64_t B) {
    const int Tmax = 1 + static_cast<int>(max_samples / sfx::kHop);
    if (pipe == kPipeSplit) {
        const size_t slice = sfx::split_slice_bytes(Tmax, c.max_pk);
        size_t n = static_cast<size_t>(split_chunk_for(slice));
        if (B > 0) n = std::min<size_t>(n, static_cast<size_t>(B));
        return sfx::kSplitHeader + slice * n;
    }
    if (pipe == kPipeStream) {
        size_t grid = static_cast<size_t>(c.sm_count) * c.stream_per_sm;
        if (B > 0) grid = std::min<size_t>(grid, static_cast<size_t>(B));
        return sfx::kWsHeader + sfx::stream_slot_bytes(Tmax, c.max_pk) * sfx::stream_slots() * grid;
    }
    size_t grid = static_cast<size_t>(c.grid_max);
    if (B > 0) grid = std::min<size_t>(grid, static_cast<size_t>(B));
    return sfx::kWsHeader + sfx::cta_scratch_bytes(Tmax, c.max_pk, pipe == kPipeFusedUmma) * grid;
}

// workspace that covers every pipeline the current mode may select for a batch of B clips (B <= 0: any batch size)
size_t mode_ws_bytes(const LaunchCtx& c, int64_t max_samples, int64_t B) {
    const int mode = g_pipeline.load();
    if (mode != 0) return pipeline_ws_bytes(c, mode, max_samples, B);
    // auto: a batch small enough for the split pipeline takes it if the workspace holds its slices, else the large-batch
    // pipeline.  Size for the split pipeline's auto range and for the large-batch pipeline.
    const int64_t bs = B > 0 ? std::min<int64_t>(B, kAutoSplitMaxB) : kAutoSplitMaxB;
    const size_t split = pipeline_ws_bytes(c, kPipeSplit, max_samples, bs);
    if (B > 0 && B <= kAutoSplitMaxBRagged) return split;      // split whatever the lengths look like
    return std::max(split, pipeline_ws_bytes(c, kAutoLarge, max_samples, B));
}

int do_extract(int device, int sr, const float* wave, int64_t row_stride, const int32_t* lengths, int64_t n_default,
               int64_t max_samples, int32_t B, int32_t n_mfcc, float* out, int64_t out_stride, void* ws,
               size_t ws_bytes, void* stream, const sfx_debug_out* dbg) {
    if (device < 0 || device >= kMaxDev) return fail(SFX_ERR_ARG, "device index out of range");
    if (B < 0 || n_mfcc < 1 || n_mfcc > sfx::kMels) return fail(SFX_ERR_ARG, "B < 0 or n_mfcc outside [1,128]");
    if (B == 0) return SFX_OK;
    if (!wave || !out || !ws) return fail(SFX_ERR_ARG, "null wave/out/workspace pointer");
    if (row_stride < 0 || out_stride < n_mfcc + 16) return fail(SFX_ERR_ARG, "bad row_stride/out_stride");
    if (max_samples < 1) return fail(SFX_ERR_ARG, "max_samples < 1");
    if (!lengths && (n_default < 1 || n_default > max_samples)) return fail(SFX_ERR_ARG, "n_default outside [1,max_samples]");
    LaunchCtx c;
    if (!snapshot(device, sr, &c)) return fail(SFX_ERR_NOT_INIT, "sfx_init_tables not called for this (device, sample rate)");
    CK(cudaSetDevice(device));
    const int Tmax = 1 + static_cast<int>(max_samples / sfx::kHop);
    sfx::Params p{};
    p.wave = wave; p.row_stride = row_stride; p.lengths = lengths; p.n_default = n_default;
    p.B = B; p.n_mfcc = n_mfcc; p.out = out; p.out_stride = out_stride;
    p.ws = static_cast<unsigned char*>(ws); p.Tmax = Tmax; p.max_samples = max_samples;
    p.aligned8 = ((reinterpret_cast<uintptr_t>(wave) & 7u) == 0 && (row_stride & 1) == 0) ? 1 : 0;
    p.tb = c.tb;
    p.max_pk = c.max_pk;
    if (dbg) p.dbg = *dbg;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t split_slice = sfx::split_slice_bytes(Tmax, c.max_pk);
    const size_t split_cap = ws_bytes > sfx::kSplitHeader
                                 ? std::min<size_t>(split_chunk_for(split_slice), (ws_bytes - sfx::kSplitHeader) / split_slice) : 0;
    const int pipe = choose_pipeline(g_pipeline.load(), B, lengths != nullptr, split_cap);
    if (pipe == kPipeSplit) {
        const size_t slice = split_slice;
        if (ws_bytes < sfx::kSplitHeader + slice)
            return fail(SFX_ERR_WORKSPACE, "workspace smaller than sfx_workspace_bytes(device, max_samples)");
        const int chunk = static_cast<int>(split_cap);
        p.cta_scratch_bytes = static_cast<long long>(slice);
        int launches = 0;
        for (int c0 = 0; c0 < B; c0 += chunk) {
            sfx::SplitParams q{};
            q.p = p;
            q.chunk0 = c0;
            q.nclips = std::min(chunk, B - c0);
            q.T_uniform = lengths ? 0 : 1 + static_cast<int>(n_default / sfx::kHop);
            const long long frames_ub = static_cast<long long>(q.nclips) * Tmax;
            const int gf = static_cast<int>(std::min<long long>((frames_ub + sfx::kWarps - 1) / sfx::kWarps, c.sm_count * c.frames_per_sm));
            const int gc = std::min(q.nclips, c.sm_count * c.clips_per_sm);
            CK(sfx::launch_split_chunk(q, std::max(gf, 1), std::max(gc, 1), dbg != nullptr, st));
            launches += 3;
        }
        g_last_launches = launches;
        return SFX_OK;
    }
    const bool stream_pipe = pipe == kPipeStream, umma = pipe == kPipeFusedUmma;
    const size_t slice = stream_pipe ? sfx::stream_slot_bytes(Tmax, c.max_pk) * sfx::stream_slots()
                                     : sfx::cta_scratch_bytes(Tmax, c.max_pk, umma);
    const int grid = static_cast<int>(std::min<int64_t>(B, stream_pipe ? c.sm_count * c.stream_per_sm : c.grid_max));
    if (ws_bytes < sfx::kWsHeader + slice * static_cast<size_t>(grid))
        return fail(SFX_ERR_WORKSPACE, "workspace smaller than sfx_workspace_bytes(device, max_samples)");
    p.cta_scratch_bytes = static_cast<long long>(stream_pipe ? sfx::stream_slot_bytes(Tmax, c.max_pk) : sfx::cta_rest_bytes(Tmax, c.max_pk));
    p.cta_p16_bytes = static_cast<long long>(sfx::cta_p16_bytes(Tmax, umma));
    p.cta_lm_bytes = static_cast<long long>(sfx::cta_lm_bytes(Tmax));
    CK(cudaMemsetAsync(ws, 0, sfx::kWsQueueBytes, st));
    g_last_launches = 1;
    if (lengths && B > grid && B <= sfx::kOrderMax) {        // ragged batch with more clips than CTAs: longest clips first
        int* order = reinterpret_cast<int*>(static_cast<unsigned char*>(ws) + sfx::kWsQueueBytes);
        CK(sfx::launch_order(lengths, B, order, st));
        p.order = order;
        g_last_launches = 2;
    }
    // Fused kernel: pin the rows phase 3 reads back (FP16 |X|^2, then log-mel, of every CTA: one block of the workspace)
    // in the persisting part of L2 for the duration of the launch.  hitRatio = the fraction of the window's lines the
    // set-aside can hold, so that persisting lines never evict each other.
    // (long clips: the block outgrows the window and nothing of it would stay resident anyway -- no pinning then, the
    // set-aside is simply unused for that launch)
    const size_t pin_bytes = static_cast<size_t>(grid) * (sfx::cta_p16_bytes(Tmax, umma) + (g_l2_mode == 2 ? sfx::cta_lm_bytes(Tmax) : 0));
    const bool l2_pin = !stream_pipe && c.l2_persist > 0 && g_l2_mode != 0 && pin_bytes <= c.l2_window;
    if (l2_pin) {
        const size_t win = pin_bytes;
        cudaStreamAttrValue av{};
        av.accessPolicyWindow.base_ptr = static_cast<unsigned char*>(ws) + sfx::kWsHeader;
        av.accessPolicyWindow.num_bytes = win;
        av.accessPolicyWindow.hitRatio = static_cast<float>(std::min(1.0, static_cast<double>(c.l2_persist) / static_cast<double>(win)));
        av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        CK(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av));
    }
    if (stream_pipe) CK(sfx::launch_stream(p, grid, dbg != nullptr, st));
    else             CK(sfx::launch_extract(p, grid, dbg != nullptr, umma, st));
    if (l2_pin) {                                            // later work on the caller's stream is not ours to steer
        cudaStreamAttrValue av{};
        av.accessPolicyWindow.num_bytes = 0;
        CK(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av));
    }
    return SFX_OK;
}

}  // namespace

template <class S>
int extract_host_impl(int device, int32_t sr, const S* host_wave, int64_t row_stride, const int32_t* host_lengths,
                      int64_t n_default, int32_t B, int32_t n_mfcc, float* host_out, int64_t out_stride,
                      int32_t chunk_clips) {
    constexpr bool kPcm = sizeof(S) == 2;
    if (device < 0 || device >= kMaxDev) return fail(SFX_ERR_ARG, "device index out of range");
    if (B < 0 || n_mfcc < 1 || n_mfcc > sfx::kMels) return fail(SFX_ERR_ARG, "B < 0 or n_mfcc outside [1,128]");
    if (B == 0) return SFX_OK;
    if (!host_wave || !host_out) return fail(SFX_ERR_ARG, "null host pointer");
    if (row_stride < 1 || out_stride < n_mfcc + 16) return fail(SFX_ERR_ARG, "bad row_stride/out_stride");
    if (!host_lengths && (n_default < 1 || n_default > row_stride)) return fail(SFX_ERR_ARG, "n_default outside [1,row_stride]");
    LaunchCtx lc;
    if (!snapshot(device, sr, &lc)) return fail(SFX_ERR_NOT_INIT, "sfx_init_tables not called for this (device, sample rate)");
    DevCtx& c = g_ctx[device];
    CK(cudaSetDevice(device));
    int64_t max_samples = n_default;
    if (host_lengths) {
        max_samples = 1;
        for (int i = 0; i < B; ++i) {
            if (host_lengths[i] <= 0 || host_lengths[i] > row_stride) return fail(SFX_ERR_BAD_CLIP, "clip length outside [1,row_stride]");
            max_samples = std::max<int64_t>(max_samples, host_lengths[i]);
        }
    }
    // rows are copied up to max_samples only (rounded to even for 8-byte alignment of every device row)
    const int64_t dev_stride = (max_samples + 1) & ~int64_t(1);
    // default chunk: ~96 MB of float rows (~380 three-second clips).  The copy engine is the bottleneck (a chunk's kernels
    // take a fifth of its H2D time), so chunks only need to be large enough to amortise launches; smaller chunks shorten the
    // un-overlapped head (first H2D) and tail (last kernel + D2H) of a call, three buffers keep the H2D queue non-empty
    int chunk = chunk_clips > 0 ? chunk_clips : static_cast<int>(std::max<int64_t>(64, (96ll << 20) / (dev_stride * 4)));
    chunk = std::min(chunk, B);
    const int out_w = n_mfcc + 16;
    const size_t need_wave = static_cast<size_t>(chunk) * dev_stride;
    const size_t need_ws = mode_ws_bytes(lc, max_samples, chunk);      // sized for the chunk, not for any batch
    const size_t need_out = static_cast<size_t>(chunk) * out_w;
    std::lock_guard<std::mutex> lk(g_host_mu);
    HostPath& hp = c.hp;
    const bool in_pinned = is_pinned(host_wave), out_pinned = is_pinned(host_out);
    if (hp.wave_elems < need_wave || hp.ws_bytes < need_ws || hp.out_elems < need_out || hp.chunk < chunk ||
        (!in_pinned && !hp.h_stage[0]) || (!out_pinned && !hp.h_out[0]) || (kPcm && !hp.d_pcm[0])) {
        free_host_path(hp);
        for (int s = 0; s < kHostStreams; ++s) {
            CK(cudaStreamCreateWithFlags(&hp.stream[s], cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&hp.ev_done[s], cudaEventDisableTiming));
            CK(cudaMalloc(&hp.d_wave[s], need_wave * 4));
            if (kPcm) CK(cudaMalloc(&hp.d_pcm[s], need_wave * 2));
            CK(cudaMalloc(&hp.d_len[s], static_cast<size_t>(chunk) * 4));
            CK(cudaMalloc(&hp.d_out[s], need_out * 4));
            CK(cudaMalloc(&hp.d_ws[s], need_ws));
            if (!in_pinned) CK(cudaMallocHost(&hp.h_stage[s], need_wave * 4));
            if (!out_pinned) CK(cudaMallocHost(&hp.h_out[s], need_out * 4));
        }
        hp.wave_elems = need_wave; hp.ws_bytes = need_ws; hp.out_elems = need_out; hp.chunk = chunk;
    }
    int nchunks = (B + chunk - 1) / chunk;
    int launches_total = 0;
    sfx::QuiesceOnError<kHostStreams> quiesce{hp.stream};
    std::vector<int> pend_c0(kHostStreams, -1), pend_nb(kHostStreams, 0);
    int64_t first_bad = -1;               // first clip whose feature row is not finite (bad length or non-finite sample)
    auto drain = [&](int s) -> int {      // copy a finished chunk's rows out of pinned staging, look for non-finite rows
        if (pend_c0[s] < 0) return SFX_OK;
        CK(cudaEventSynchronize(hp.ev_done[s]));
        for (int i = 0; i < pend_nb[s]; ++i) {
            float* dst = host_out + static_cast<int64_t>(pend_c0[s] + i) * out_stride;
            if (!out_pinned) std::memcpy(dst, hp.h_out[s] + static_cast<size_t>(i) * out_w, sizeof(float) * out_w);
            // a NaN / Inf sample reaches every frame-mean it touches (rms and centroid at least): one look at the row's
            // four descriptors finds it without a pass over the waveform on the host
            const float chk = dst[n_mfcc + 12] + dst[n_mfcc + 13] + dst[n_mfcc + 14] + dst[n_mfcc + 15] + dst[0];
            if (!std::isfinite(chk) && (first_bad < 0 || pend_c0[s] + i < first_bad)) first_bad = pend_c0[s] + i;
        }
        pend_c0[s] = -1;
        return SFX_OK;
    };
    for (int ci = 0; ci < nchunks; ++ci) {
        const int s = ci % kHostStreams;
        const int c0 = ci * chunk, nb = std::min(chunk, B - c0);
        int rc = drain(s);
        if (rc) return rc;
        cudaStream_t st = hp.stream[s];
        const S* src = host_wave + static_cast<int64_t>(c0) * row_stride;
        const size_t row_bytes = static_cast<size_t>(max_samples) * sizeof(S);
        void* d_in = kPcm ? static_cast<void*>(hp.d_pcm[s]) : static_cast<void*>(hp.d_wave[s]);
        if (in_pinned) {
            CK(cudaMemcpy2DAsync(d_in, dev_stride * sizeof(S), src, row_stride * sizeof(S), row_bytes, nb, cudaMemcpyHostToDevice, st));
        } else {
            S* stage = reinterpret_cast<S*>(hp.h_stage[s]);
            for (int i = 0; i < nb; ++i)
                std::memcpy(stage + static_cast<size_t>(i) * dev_stride, src + static_cast<int64_t>(i) * row_stride, row_bytes);
            CK(cudaMemcpyAsync(d_in, stage, static_cast<size_t>(nb) * dev_stride * sizeof(S), cudaMemcpyHostToDevice, st));
        }
        if (kPcm) {
            const long long pairs = static_cast<long long>(nb) * dev_stride / 2;
            const int grid = static_cast<int>(std::min<long long>((pairs + 255) / 256, 148ll * 16));
            pcm16_to_f32_kernel<<<grid, 256, 0, st>>>(hp.d_pcm[s], hp.d_wave[s], pairs);
            CK(cudaGetLastError());
            ++launches_total;
        }
        const int32_t* dlen = nullptr;
        if (host_lengths) {
            CK(cudaMemcpyAsync(hp.d_len[s], host_lengths + c0, static_cast<size_t>(nb) * 4, cudaMemcpyHostToDevice, st));
            dlen = hp.d_len[s];
        }
        rc = do_extract(device, sr, hp.d_wave[s], dev_stride, dlen, n_default, max_samples, nb, n_mfcc, hp.d_out[s], out_w,
                        hp.d_ws[s], hp.ws_bytes, st, nullptr);
        if (rc) return rc;
        launches_total += g_last_launches;
        if (out_pinned) {
            CK(cudaMemcpy2DAsync(host_out + static_cast<int64_t>(c0) * out_stride, out_stride * 4, hp.d_out[s], out_w * 4,
                                 static_cast<size_t>(out_w) * 4, nb, cudaMemcpyDeviceToHost, st));
        } else {
            CK(cudaMemcpyAsync(hp.h_out[s], hp.d_out[s], static_cast<size_t>(nb) * out_w * 4, cudaMemcpyDeviceToHost, st));
        }
        CK(cudaEventRecord(hp.ev_done[s], st));
        pend_c0[s] = c0; pend_nb[s] = nb;
    }
    for (int s = 0; s < kHostStreams; ++s) {
        int rc = drain(s);
        if (rc) return rc;
    }
    quiesce.armed = false;
    g_last_launches = launches_total;
    if (first_bad >= 0)
        return fail(SFX_ERR_BAD_CLIP, "clip " + std::to_string(first_bad) + ": non-finite sample (its feature row, like that of every "
                                      "other such clip, is NaN; all rows were delivered)");
    return SFX_OK;
}


extern "C" {

int sfx_abi_version(void) { return SFX_ABI_VERSION; }
const char* sfx_last_error(void) { return g_err.c_str(); }

int sfx_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); g_err = cudaGetErrorString(e); return e == cudaErrorNoDevice ? 0 : SFX_ERR_CUDA; }
    return n;
}

int sfx_launches_per_extract(void) { return g_last_launches; }

int sfx_set_pipeline(int mode) {
    if (mode < 0 || mode > 4) return fail(SFX_ERR_ARG, "pipeline mode must be 0 (auto), 1 (fused), 2 (split), 3 (stream) or 4 (fused_umma)");
    g_pipeline.store(mode);
    return SFX_OK;
}

int sfx_release(int device) {
    if (device < 0 || device >= kMaxDev) return fail(SFX_ERR_ARG, "device index out of range");
    sfx_frontend_release(device);
    std::lock_guard<std::mutex> lh(g_host_mu);
    std::lock_guard<std::mutex> lk(g_tab_mu);
    DevCtx& c = g_ctx[device];
    if (!c.ready && c.allocs.empty()) return SFX_OK;
    cudaSetDevice(device);
    free_host_path(c.hp);
    for (void* d : c.allocs) cudaFree(d);
    c = DevCtx{};
    return SFX_OK;
}

int sfx_init_tables(int device, const sfx_tables_host* t) {
    if (device < 0 || device >= kMaxDev) return fail(SFX_ERR_ARG, "device index out of range");
    if (!t || !t->hann || !t->tw1 || !t->tw2 || !t->mel_ab || !t->mel_mask || !t->mel_src || !t->chroma16 || !t->chroma_ny || !t->dct ||
        !t->edges || !t->chroma_frag || !t->chroma_umma || t->sr <= 0)
        return fail(SFX_ERR_ARG, "null table pointer or bad sizes");
    if (t->mel_ps < 3 || (t->mel_ps & 1) == 0 || sfx::kPartOff + 32 * t->mel_ps + 1 > sfx::kExFloats)
        return fail(SFX_ERR_ARG, "mel_ps must be odd and fit the warp tile");
    if (t->pip_kmin < 1 || t->pip_kmax > sfx::kBins - 2 || t->pip_kmax < t->pip_kmin)
        return fail(SFX_ERR_ARG, "piptrack bin range outside [1,1023]");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device >= ndev) return fail(SFX_ERR_CUDA, "no such CUDA device");
    std::lock_guard<std::mutex> lk(g_tab_mu);
    DevCtx& c = g_ctx[device];
    if (c.ready && find_set(c, t->sr)) return SFX_OK;        // idempotent per (device, sr)
    CK(cudaSetDevice(device));
    cudaDeviceProp prop{};
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(SFX_ERR_CUDA, "libsfx_b200 is built for sm_100a only");
    c.sm_count = prop.multiProcessorCount;
    if (g_l2_mode != 0 && c.l2_persist == 0 && prop.persistingL2CacheMaxSize > 0) {
        // set aside the most L2 the device allows for persisting accesses (79 of 126 MB on a B200); device-wide, once
        if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, static_cast<size_t>(prop.persistingL2CacheMaxSize)) == cudaSuccess) {
            size_t got = 0;
            if (cudaDeviceGetLimit(&got, cudaLimitPersistingL2CacheSize) == cudaSuccess) c.l2_persist = got;
            c.l2_window = static_cast<size_t>(prop.accessPolicyMaxWindowSize);
        } else {
            cudaGetLastError();
        }
    }
    TableSet set;
    set.sr = t->sr;
    int rc;
    const float* f = nullptr;
    if ((rc = upload(c, t->hann, 2048, &f))) return rc;
    set.tb.hann = reinterpret_cast<const float2*>(f);
    if ((rc = upload(c, t->tw1, 2048, &f))) return rc;
    set.tb.tw1 = reinterpret_cast<const float2*>(f);
    if ((rc = upload(c, t->tw2, 2048, &f))) return rc;
    set.tb.tw2 = reinterpret_cast<const float2*>(f);
    if ((rc = upload(c, t->mel_ab, 33 * 32 * 2, &f))) return rc;
    set.tb.mel_ab = reinterpret_cast<const float2*>(f);
    if ((rc = upload(c, t->mel_mask, 32, &set.tb.mel_mask))) return rc;
    if ((rc = upload(c, t->mel_src, 128 * 3, &set.tb.mel_src))) return rc;
    if ((rc = upload(c, t->chroma16, static_cast<size_t>(sfx::kTunings) * 2 * sfx::kChroma * sfx::kP16Stride, &set.tb.chroma16))) return rc;
    if ((rc = upload(c, t->chroma_ny, static_cast<size_t>(sfx::kTunings) * sfx::kChroma, &set.tb.chroma_ny))) return rc;
    {
        const uint32_t* frag = nullptr;
        if ((rc = upload(c, t->chroma_frag, static_cast<size_t>(sfx::kTunings) * 32 * 2 * 2 * 32 * 4, &frag))) return rc;
        set.tb.chroma_frag = reinterpret_cast<const uint4*>(frag);
    }
    if ((rc = upload(c, t->chroma_umma, static_cast<size_t>(sfx::kTunings) * 16 * 4096, &set.tb.chroma_umma))) return rc;
    std::vector<double> dctT(static_cast<size_t>(sfx::kMels) * sfx::kMels);
    for (int k = 0; k < sfx::kMels; ++k)
        for (int m = 0; m < sfx::kMels; ++m) dctT[static_cast<size_t>(m) * sfx::kMels + k] = t->dct[static_cast<size_t>(k) * sfx::kMels + m];
    if ((rc = upload(c, dctT.data(), dctT.size(), &set.tb.dctT))) return rc;
    if ((rc = upload(c, t->edges, sfx::kTunings + 1, &set.tb.edges))) return rc;
    set.tb.mel_ps = t->mel_ps; set.tb.mel_flush32 = t->mel_flush32;
    set.tb.sr = t->sr; set.tb.kmin = t->pip_kmin; set.tb.kmax = t->pip_kmax;
    CK(sfx::configure_kernels(&c.blocks_per_sm));
    if (c.blocks_per_sm < 1) return fail(SFX_ERR_CUDA, "kernel does not fit on an SM");
    if (const char* e = std::getenv("SFX_BLOCKS_PER_SM")) {      // experiment: override the occupancy calculator's answer
        const int v = std::atoi(e);
        if (v >= 1 && v <= 4) c.blocks_per_sm = v;
    }
    c.grid_max = c.sm_count * c.blocks_per_sm;
    CK(sfx::configure_split(&c.frames_per_sm, &c.clips_per_sm));
    if (c.frames_per_sm < 1 || c.clips_per_sm < 1) return fail(SFX_ERR_CUDA, "split kernels do not fit on an SM");
    CK(sfx::configure_stream(&c.stream_per_sm));
    if (c.stream_per_sm < 1) return fail(SFX_ERR_CUDA, "stream kernel does not fit on an SM");
    c.max_pk = std::max(c.max_pk, (((t->pip_kmax - t->pip_kmin + 2) / 2) + 3) & ~3);
    c.sets.push_back(set);
    c.ready = true;
    return SFX_OK;
}

size_t sfx_workspace_bytes_batch(int device, int64_t max_samples, int64_t B) {
    LaunchCtx c;
    bool ok = device >= 0 && device < kMaxDev && max_samples >= 1;
    if (ok) {
        std::lock_guard<std::mutex> lk(g_tab_mu);
        const DevCtx& d = g_ctx[device];
        ok = d.ready;
        c.sm_count = d.sm_count; c.grid_max = d.grid_max; c.stream_per_sm = d.stream_per_sm; c.max_pk = d.max_pk;
    }
    if (!ok) {
        g_err = "sfx_workspace_bytes: bad device/max_samples or tables not initialised";
        return 0;
    }
    return mode_ws_bytes(c, max_samples, B);
}

size_t sfx_workspace_bytes(int device, int64_t max_samples) { return sfx_workspace_bytes_batch(device, max_samples, 0); }

int sfx_extract(int device, int32_t sr, const float* wave, int64_t row_stride, const int32_t* lengths, int64_t n_default,
                int64_t max_samples, int32_t B, int32_t n_mfcc, float* out, int64_t out_stride, void* workspace,
                size_t workspace_bytes, void* stream) {
    return do_extract(device, sr, wave, row_stride, lengths, n_default, max_samples, B, n_mfcc, out, out_stride, workspace,
                      workspace_bytes, stream, nullptr);
}

int sfx_extract_debug(int device, int32_t sr, const float* wave, int64_t row_stride, const int32_t* lengths, int64_t n_default,
                      int64_t max_samples, int32_t B, int32_t n_mfcc, float* out, int64_t out_stride, void* workspace,
                      size_t workspace_bytes, void* stream, const sfx_debug_out* dbg) {
    if (!dbg) return fail(SFX_ERR_ARG, "null dbg");
    return do_extract(device, sr, wave, row_stride, lengths, n_default, max_samples, B, n_mfcc, out, out_stride, workspace,
                      workspace_bytes, stream, dbg);
}

int sfx_extract_host(int device, int32_t sr, const float* host_wave, int64_t row_stride, const int32_t* host_lengths,
                     int64_t n_default, int32_t B, int32_t n_mfcc, float* host_out, int64_t out_stride,
                     int32_t chunk_clips) {
    return extract_host_impl<float>(device, sr, host_wave, row_stride, host_lengths, n_default, B, n_mfcc, host_out, out_stride,
                                    chunk_clips);
}

int sfx_extract_host_pcm16(int device, int32_t sr, const int16_t* host_pcm, int64_t row_stride, const int32_t* host_lengths,
                           int64_t n_default, int32_t B, int32_t n_mfcc, float* host_out, int64_t out_stride,
                           int32_t chunk_clips) {
    return extract_host_impl<int16_t>(device, sr, host_pcm, row_stride, host_lengths, n_default, B, n_mfcc, host_out, out_stride,
                                      chunk_clips);
}

}  // extern "C"
