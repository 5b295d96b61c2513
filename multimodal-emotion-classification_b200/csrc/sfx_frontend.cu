// sfx_frontend.cu -- scope row f3: the load_audio front-end (reference preprocessing/audio_preprocessing.py:12-19) for
// 16-bit PCM on the device: x / 32768 as soundfile does, mono mean, polyphase resampling to the extractor's rate, zero
// pad / trim to sr * duration, then the feature extractor -- all from the raw PCM rows of the WAV files, in one call.
//
// The resampler is scipy.signal.resample_poly (the resampler of this package's load_audio; librosa's soxr_hq cannot be
// restated or checked here): y[m] = sum_i x[i] * h[(m + n_pre_remove) * down - i * up] over the taps that
// exist, accumulated in float64 in ascending i with separate multiply and add -- the order and rounding of scipy's
// _upfirdn_apply -- and rounded once to float32.  The output is bit-identical to the host path.
#include <algorithm>
#include <mutex>
#include <type_traits>
#include <string>
#include <vector>

#include "sfx_internal.h"

namespace {

constexpr int kMaxDev = 64;
constexpr int kStreams = 2;

struct Filter {                 // device copy of one (up, down) filter
    int up = 0, down = 0, n_taps = 0, n_pre_remove = 0;
    int per_phase = 0;          // K = ceil(n_taps / up)
    double* d_taps = nullptr;   // [K][up], output-major: [k][r] = tap r' + k * up of the phase r' of outputs m = r (mod up)
};

struct FrontPath {
    cudaStream_t stream[kStreams] = {nullptr, nullptr};
    cudaEvent_t ev_done[kStreams] = {nullptr, nullptr};
    int16_t* d_pcm[kStreams] = {nullptr, nullptr};
    int32_t* d_frames[kStreams] = {nullptr, nullptr};
    float* d_wave[kStreams] = {nullptr, nullptr};
    float* d_out[kStreams] = {nullptr, nullptr};
    void* d_ws[kStreams] = {nullptr, nullptr};
    int16_t* h_stage[kStreams] = {nullptr, nullptr};
    float* h_out[kStreams] = {nullptr, nullptr};
    size_t pcm_elems = 0, wave_elems = 0, out_elems = 0, ws_bytes = 0;
    int chunk = 0;
    std::vector<Filter> filters;
};

FrontPath g_front[kMaxDev];
std::mutex g_front_mu;
thread_local std::string g_front_err;

int ffail(int code, const std::string& msg) { g_front_err = msg; return code; }
#define FCK(call)                                                                              \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) return ffail(SFX_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

void free_front(FrontPath& fp, bool filters_too) {
    for (int s = 0; s < kStreams; ++s) {
        if (fp.d_pcm[s]) cudaFree(fp.d_pcm[s]);
        if (fp.d_frames[s]) cudaFree(fp.d_frames[s]);
        if (fp.d_wave[s]) cudaFree(fp.d_wave[s]);
        if (fp.d_out[s]) cudaFree(fp.d_out[s]);
        if (fp.d_ws[s]) cudaFree(fp.d_ws[s]);
        if (fp.h_stage[s]) cudaFreeHost(fp.h_stage[s]);
        if (fp.h_out[s]) cudaFreeHost(fp.h_out[s]);
        if (fp.ev_done[s]) cudaEventDestroy(fp.ev_done[s]);
        if (fp.stream[s]) cudaStreamDestroy(fp.stream[s]);
        fp.d_pcm[s] = nullptr; fp.d_frames[s] = nullptr; fp.d_wave[s] = nullptr; fp.d_out[s] = nullptr; fp.d_ws[s] = nullptr;
        fp.h_stage[s] = nullptr; fp.h_out[s] = nullptr; fp.ev_done[s] = nullptr; fp.stream[s] = nullptr;
    }
    fp.pcm_elems = fp.wave_elems = fp.out_elems = fp.ws_bytes = 0;
    fp.chunk = 0;
    if (filters_too) {
        for (Filter& f : fp.filters) cudaFree(f.d_taps);
        fp.filters.clear();
    }
}

bool pinned(const void* p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

// One thread per output sample, a block = 256 consecutive outputs of one clip.  Output m uses the taps of phase
// ((m + n_pre_remove) * down) % up, which depends on m % up only, so the filter is stored "output-major":
// tab[k][r] = tap number k of the phase that outputs m = r (mod up) use.  For a given k the 32 lanes of a warp then read 32
// consecutive doubles (and nearly consecutive samples), whatever `up` is -- 147 for 48 kHz -> 22.05 kHz or 3675 for
// TESS's 24 414 Hz.  Sample i = q - k pairs with tab[k][r] (q = (m + n_pre_remove) * down / up); k runs downwards so that
// the sum runs over ascending i like scipy's.  Taps past the end of the filter are stored as zeros: adding x * 0 does
// not change the float64 sum, so every thread runs the same K steps.
// Exactness: a sample is k * 2^-15 (mono) or (k_l + k_r) * 2^-16 (stereo mean, exact in float32), so summing
// round(k * h) and scaling the total by the power of two reproduces, rounding for rounding, scipy's sum of round(x * h);
// the integer converts to float64 in one instruction.
// kWide = false: every index product fits 31 bits (checked on the host), so the divisions are 32-bit.
template <bool kWide>
__global__ void __launch_bounds__(256) sfx_resample_pcm16_kernel(
    const int16_t* __restrict__ pcm, const long long row_stride, const int channels, const int32_t* __restrict__ frames,
    const int frames_default, const int up, const int down, const double* __restrict__ tab, const int per_phase,
    const int n_pre_remove, float* __restrict__ wave, const long long wave_stride, const int n_target) {
    const int b = blockIdx.y;
    const int m = blockIdx.x * 256 + threadIdx.x;
    if (m >= n_target) return;
    const int n_in = frames ? frames[b] : frames_default;
    const int16_t* x = pcm + static_cast<long long>(b) * row_stride;
    float* y = wave + static_cast<long long>(b) * wave_stride;
    const double scale = channels == 1 ? 1.0 / 32768.0 : 1.0 / 65536.0;
    auto ksum = [&](int i) -> int {                           // integer numerator of the (mono-mixed) sample
        if (channels == 1) return static_cast<int>(x[i]);
        const int v = reinterpret_cast<const int*>(x)[i];     // one 32-bit load per stereo frame (rows are 4-byte aligned)
        return static_cast<int>(static_cast<short>(v & 0xffff)) + (v >> 16);
    };
    if (n_in <= 0) { y[m] = 0.0f; return; }
    if (up == down) {                                   // native rate: no filter (resample_poly returns a copy)
        y[m] = m < n_in ? static_cast<float>(static_cast<double>(ksum(m)) * scale) : 0.0f;
        return;
    }
    using idx_t = std::conditional_t<kWide, long long, int>;
    const idx_t n_out = (static_cast<idx_t>(n_in) * up + down - 1) / down;
    if (m >= n_out) { y[m] = 0.0f; return; }
    const int q = static_cast<int>((static_cast<idx_t>(m) + n_pre_remove) * down / up);
    const double* t = tab + static_cast<size_t>(per_phase - 1) * up + m % up;
    double acc = 0.0;
    const int i0 = q - (per_phase - 1);                 // sample of the first step
    if (i0 >= 0 && q < n_in) {                          // interior output: all K samples exist, no per-step checks
        if (channels == 1) {
            const int16_t* xs = x + i0;
#pragma unroll 4
            for (int j = 0; j < per_phase; ++j)
                acc = __dadd_rn(acc, __dmul_rn(static_cast<double>(static_cast<int>(xs[j])), __ldg(t - static_cast<size_t>(j) * up)));
        } else {
            const int* xs = reinterpret_cast<const int*>(x) + i0;
#pragma unroll 4
            for (int j = 0; j < per_phase; ++j) {
                const int v = xs[j];
                const int ks = static_cast<int>(static_cast<short>(v & 0xffff)) + (v >> 16);
                acc = __dadd_rn(acc, __dmul_rn(static_cast<double>(ks), __ldg(t - static_cast<size_t>(j) * up)));
            }
        }
    } else {
        for (int j = 0; j < per_phase; ++j, t -= up) {
            const int i = i0 + j;
            if (i >= 0 && i < n_in) acc = __dadd_rn(acc, __dmul_rn(static_cast<double>(ksum(i)), __ldg(t)));
        }
    }
    y[m] = static_cast<float>(acc * scale);
}

}  // namespace

extern "C" {

const char* sfx_frontend_last_error(void) { return g_front_err.c_str(); }

int sfx_frontend_release(int device) {
    if (device < 0 || device >= kMaxDev) return ffail(SFX_ERR_ARG, "device index out of range");
    std::lock_guard<std::mutex> lk(g_front_mu);
    cudaSetDevice(device);
    free_front(g_front[device], true);
    return SFX_OK;
}

int sfx_preprocess_host_pcm16(int device, int32_t sr, const sfx_resampler_host* rs, const int16_t* host_pcm,
                              int64_t row_stride, int32_t channels, const int32_t* host_frames, int64_t frames_default,
                              int64_t n_target, int32_t B, int32_t n_mfcc, float* host_out, int64_t out_stride,
                              int32_t chunk_clips) {
    if (device < 0 || device >= kMaxDev) return ffail(SFX_ERR_ARG, "device index out of range");
    if (B < 0 || n_mfcc < 1 || n_mfcc > sfx::kMels) return ffail(SFX_ERR_ARG, "B < 0 or n_mfcc outside [1,128]");
    if (B == 0) return SFX_OK;
    if (!host_pcm || !host_out) return ffail(SFX_ERR_ARG, "null host pointer");
    if (channels != 1 && channels != 2) return ffail(SFX_ERR_ARG, "channels must be 1 or 2");
    if (n_target < 1 || out_stride < n_mfcc + 16) return ffail(SFX_ERR_ARG, "bad n_target/out_stride");
    const int up = rs ? rs->up : 1, down = rs ? rs->down : 1;
    if (rs && (up < 1 || down < 1 || (up != down && (!rs->taps || rs->n_taps < 1 || rs->n_pre_remove < 0))))
        return ffail(SFX_ERR_ARG, "bad resampler description");
    int64_t max_frames = frames_default;
    if (host_frames) {
        max_frames = 0;
        for (int i = 0; i < B; ++i) {
            if (host_frames[i] < 0) return ffail(SFX_ERR_BAD_CLIP, "negative frame count");
            max_frames = std::max<int64_t>(max_frames, host_frames[i]);
        }
    }
    if (max_frames < 1 || max_frames * channels > row_stride) return ffail(SFX_ERR_ARG, "frames outside [1, row_stride / channels]");
    int ndev = 0;
    FCK(cudaGetDeviceCount(&ndev));
    if (device >= ndev) return ffail(SFX_ERR_CUDA, "no such CUDA device");
    FCK(cudaSetDevice(device));
    const int64_t pcm_stride = (max_frames * channels + 1) & ~int64_t(1);         // int16 elements per device row
    const int64_t wave_stride = (n_target + 1) & ~int64_t(1);
    int chunk = chunk_clips > 0 ? chunk_clips
                                : static_cast<int>(std::max<int64_t>(64, (256ll << 20) / std::max(pcm_stride * 2, wave_stride * 4)));
    chunk = std::min(chunk, B);
    const int out_w = n_mfcc + 16;
    const size_t need_pcm = static_cast<size_t>(chunk) * pcm_stride, need_wave = static_cast<size_t>(chunk) * wave_stride;
    const size_t need_out = static_cast<size_t>(chunk) * out_w;
    const size_t need_ws = sfx_workspace_bytes_batch(device, n_target, chunk);       // sized for the chunk, not for any batch
    if (need_ws == 0) return ffail(SFX_ERR_NOT_INIT, "sfx_init_tables not called for this (device, sample rate)");

    std::lock_guard<std::mutex> lk(g_front_mu);
    FrontPath& fp = g_front[device];
    const bool in_pinned = pinned(host_pcm), out_pinned = pinned(host_out);
    if (fp.pcm_elems < need_pcm || fp.wave_elems < need_wave || fp.out_elems < need_out || fp.ws_bytes < need_ws || fp.chunk < chunk ||
        (!in_pinned && !fp.h_stage[0]) || (!out_pinned && !fp.h_out[0])) {
        free_front(fp, false);
        for (int s = 0; s < kStreams; ++s) {
            FCK(cudaStreamCreateWithFlags(&fp.stream[s], cudaStreamNonBlocking));
            FCK(cudaEventCreateWithFlags(&fp.ev_done[s], cudaEventDisableTiming));
            FCK(cudaMalloc(&fp.d_pcm[s], need_pcm * 2));
            FCK(cudaMalloc(&fp.d_frames[s], static_cast<size_t>(chunk) * 4));
            FCK(cudaMalloc(&fp.d_wave[s], need_wave * 4));
            FCK(cudaMalloc(&fp.d_out[s], need_out * 4));
            FCK(cudaMalloc(&fp.d_ws[s], need_ws));
            if (!in_pinned) FCK(cudaMallocHost(&fp.h_stage[s], need_pcm * 2));
            if (!out_pinned) FCK(cudaMallocHost(&fp.h_out[s], need_out * 4));
        }
        fp.pcm_elems = need_pcm; fp.wave_elems = need_wave; fp.out_elems = need_out; fp.ws_bytes = need_ws; fp.chunk = chunk;
    }
    const Filter* flt = nullptr;
    if (up != down) {
        for (const Filter& f : fp.filters)
            if (f.up == up && f.down == down && f.n_taps == rs->n_taps && f.n_pre_remove == rs->n_pre_remove) flt = &f;
        if (!flt) {
            Filter f;
            f.up = up; f.down = down; f.n_taps = rs->n_taps; f.n_pre_remove = rs->n_pre_remove;
            f.per_phase = (f.n_taps + up - 1) / up;
            std::vector<double> poly(static_cast<size_t>(up) * f.per_phase, 0.0);
            for (int r = 0; r < up; ++r) {
                const long long phase = (static_cast<long long>(r) + f.n_pre_remove) * down % up;
                for (int k = 0; k < f.per_phase; ++k) {
                    const long long tap = phase + static_cast<long long>(k) * up;
                    if (tap < f.n_taps) poly[static_cast<size_t>(k) * up + r] = rs->taps[tap];
                }
            }
            FCK(cudaMalloc(&f.d_taps, sizeof(double) * poly.size()));
            FCK(cudaMemcpy(f.d_taps, poly.data(), sizeof(double) * poly.size(), cudaMemcpyHostToDevice));
            fp.filters.push_back(f);
            flt = &fp.filters.back();
        }
    }
    const int nchunks = (B + chunk - 1) / chunk;
    std::vector<int> pend_c0(kStreams, -1), pend_nb(kStreams, 0);
    auto drain = [&](int s) -> int {
        if (pend_c0[s] < 0) return SFX_OK;
        FCK(cudaEventSynchronize(fp.ev_done[s]));
        if (!out_pinned)
            for (int i = 0; i < pend_nb[s]; ++i)
                std::copy_n(fp.h_out[s] + static_cast<size_t>(i) * out_w, out_w, host_out + static_cast<int64_t>(pend_c0[s] + i) * out_stride);
        pend_c0[s] = -1;
        return SFX_OK;
    };
    const size_t row_bytes = static_cast<size_t>(max_frames) * channels * 2;
    sfx::QuiesceOnError<kStreams> quiesce{fp.stream};
    for (int ci = 0; ci < nchunks; ++ci) {
        const int s = ci % kStreams;
        const int c0 = ci * chunk, nb = std::min(chunk, B - c0);
        int rc = drain(s);
        if (rc) return rc;
        cudaStream_t st = fp.stream[s];
        const int16_t* src = host_pcm + static_cast<int64_t>(c0) * row_stride;
        if (in_pinned) {
            FCK(cudaMemcpy2DAsync(fp.d_pcm[s], pcm_stride * 2, src, row_stride * 2, row_bytes, nb, cudaMemcpyHostToDevice, st));
        } else {
            for (int i = 0; i < nb; ++i)
                std::copy_n(src + static_cast<int64_t>(i) * row_stride, max_frames * channels, fp.h_stage[s] + static_cast<size_t>(i) * pcm_stride);
            FCK(cudaMemcpyAsync(fp.d_pcm[s], fp.h_stage[s], static_cast<size_t>(nb) * pcm_stride * 2, cudaMemcpyHostToDevice, st));
        }
        const int32_t* dfr = nullptr;
        if (host_frames) {
            FCK(cudaMemcpyAsync(fp.d_frames[s], host_frames + c0, static_cast<size_t>(nb) * 4, cudaMemcpyHostToDevice, st));
            dfr = fp.d_frames[s];
        }
        const double* d_taps = flt ? flt->d_taps : nullptr;
        const int per_phase = flt ? flt->per_phase : 0, npr = flt ? flt->n_pre_remove : 0;
        // 32-bit index arithmetic when frames * up and (n_target + n_pre_remove) * down stay below 2^31
        const bool wide = flt && ((max_frames + 2) * up + down >= (1ll << 31) || (n_target + npr + 2) * static_cast<long long>(down) >= (1ll << 31));
        const dim3 grid(static_cast<unsigned>((n_target + 255) / 256), static_cast<unsigned>(nb));
        if (wide)
            sfx_resample_pcm16_kernel<true><<<grid, 256, 0, st>>>(fp.d_pcm[s], pcm_stride, channels, dfr, static_cast<int>(frames_default),
                                                                  up, down, d_taps, per_phase, npr, fp.d_wave[s], wave_stride,
                                                                  static_cast<int>(n_target));
        else
            sfx_resample_pcm16_kernel<false><<<grid, 256, 0, st>>>(fp.d_pcm[s], pcm_stride, channels, dfr, static_cast<int>(frames_default),
                                                                   up, down, d_taps, per_phase, npr, fp.d_wave[s], wave_stride,
                                                                   static_cast<int>(n_target));
        FCK(cudaGetLastError());
        rc = sfx_extract(device, sr, fp.d_wave[s], wave_stride, nullptr, n_target, n_target, nb, n_mfcc, fp.d_out[s], out_w, fp.d_ws[s],
                         fp.ws_bytes, st);
        if (rc) return ffail(rc, sfx_last_error());
        if (out_pinned) {
            FCK(cudaMemcpy2DAsync(host_out + static_cast<int64_t>(c0) * out_stride, out_stride * 4, fp.d_out[s], out_w * 4,
                                  static_cast<size_t>(out_w) * 4, nb, cudaMemcpyDeviceToHost, st));
        } else {
            FCK(cudaMemcpyAsync(fp.h_out[s], fp.d_out[s], static_cast<size_t>(nb) * out_w * 4, cudaMemcpyDeviceToHost, st));
        }
        FCK(cudaEventRecord(fp.ev_done[s], st));
        pend_c0[s] = c0; pend_nb[s] = nb;
    }
    for (int s = 0; s < kStreams; ++s) {
        int rc = drain(s);
        if (rc) return rc;
    }
    quiesce.armed = false;
    return SFX_OK;
}

}  // extern "C"
