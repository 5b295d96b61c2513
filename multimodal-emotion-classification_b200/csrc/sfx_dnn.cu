// sfx_dnn.cu -- device-resident scaler + speech DNN forward (scope row f1): the consumer of the 56-d features in
// the reference's inference/speech_inference.py:66-76,85-105 (architecture: model_training/train_speech_model.py:55-90).
// FP32 throughout: StandardScaler -> [Dense + BatchNorm(eps) + ReLU] x (L-1) -> Dense + softmax, with the last hidden
// activation (Keras layers[-3]) returned as the fusion feature tap.  0.93 MFLOP per clip (< 10 % of the extraction),
// so the dense layers are a plain 64x64x16 shared-memory tiled SIMT GEMM with the per-channel affine + ReLU fused in.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/sfx.h"

namespace {

constexpr int kMaxLayers = 8;

struct Dnn {
    int device = 0;
    int n_layers = 0;
    int dims[kMaxLayers + 1] = {0};
    float* kernel[kMaxLayers] = {nullptr};   // [in][out]
    float* scale[kMaxLayers] = {nullptr};    // per output channel: BN folded gamma / sqrt(var + eps) (1 for the last layer)
    float* shift[kMaxLayers] = {nullptr};    // (bias - mean) * scale + beta   (bias for the last layer)
    float* pre_mean = nullptr;               // scaler mean_  [dims[0]]
    float* pre_inv = nullptr;                // 1 / scaler scale_
    std::vector<void*> allocs;
};

thread_local std::string g_dnn_err;
int dfail(int code, const std::string& m) { g_dnn_err = m; return code; }

// C[M x N] = act((pre(A)[M x K] . W[K x N]) * scale[n] + shift[n]);  pre(a)[k] = (a - mean[k]) * inv[k] when mean != null
template <bool kRelu>
__global__ void __launch_bounds__(256) dense_affine_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ W,
                                                          const float* __restrict__ scale, const float* __restrict__ shift,
                                                          const float* __restrict__ pre_mean, const float* __restrict__ pre_inv,
                                                          float* __restrict__ C, long long ldc, int M, int N, int K) {
    __shared__ float As[16][64 + 4];
    __shared__ float Ws[16][64 + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += 16) {
        for (int i = threadIdx.x; i < 64 * 16; i += 256) {
            const int r = i >> 4, kk = i & 15;                       // A tile: 64 rows x 16 k
            const int m = m0 + r, k = k0 + kk;
            float v = 0.0f;
            if (m < M && k < K) {
                v = A[m * lda + k];
                if (pre_mean) v = (v - pre_mean[k]) * pre_inv[k];
            }
            As[kk][r] = v;
            const int kr = i >> 6, c = i & 63;                       // W tile: 16 k x 64 cols
            const int kw = k0 + kr, n = n0 + c;
            Ws[kr][c] = (kw < K && n < N) ? W[static_cast<long long>(kw) * N + n] : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Ws[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = fmaf(acc[i][j], scale[n], shift[n]);
            if (kRelu) v = fmaxf(v, 0.0f);
            C[m * ldc + n] = v;
        }
    }
}

// in-place softmax over rows of at most 32 logits (one warp per row)
__global__ void softmax_rows_kernel(float* __restrict__ X, long long ld, int M, int N) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= M) return;
    float v = lane < N ? X[row * ld + lane] : -INFINITY;
    float mx = v;
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float e = lane < N ? expf(v - mx) : 0.0f;
    float s = e;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane < N) X[row * ld + lane] = e / s;
}

template <class T>
int up(Dnn* d, const T* host, size_t n, T** dev) {
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, n * sizeof(T));
    if (e != cudaSuccess) return dfail(SFX_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    d->allocs.push_back(p);
    e = cudaMemcpy(p, host, n * sizeof(T), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return dfail(SFX_ERR_CUDA, std::string("cudaMemcpy: ") + cudaGetErrorString(e));
    *dev = static_cast<T*>(p);
    return SFX_OK;
}

}  // namespace

extern "C" {

const char* sfx_dnn_last_error(void) { return g_dnn_err.c_str(); }

int sfx_dnn_destroy(void* handle) {
    Dnn* d = static_cast<Dnn*>(handle);
    if (!d) return SFX_OK;
    cudaSetDevice(d->device);
    for (void* p : d->allocs) cudaFree(p);
    delete d;
    return SFX_OK;
}

int sfx_dnn_create(int device, const sfx_dnn_host* h, void** handle) {
    if (!h || !handle) return dfail(SFX_ERR_ARG, "null argument");
    if (h->n_layers < 1 || h->n_layers > kMaxLayers || !h->dims || !h->kernel || !h->bias)
        return dfail(SFX_ERR_ARG, "n_layers outside [1,8] or null arrays");
    if (h->dims[h->n_layers] > 32) return dfail(SFX_ERR_ARG, "softmax width > 32 not supported");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); return dfail(SFX_ERR_CUDA, "no such CUDA device"); }
    cudaSetDevice(device);
    Dnn* d = new Dnn();
    d->device = device;
    d->n_layers = h->n_layers;
    for (int i = 0; i <= h->n_layers; ++i) d->dims[i] = h->dims[i];
    int rc = SFX_OK;
    for (int l = 0; l < h->n_layers && rc == SFX_OK; ++l) {
        const int in = h->dims[l], out = h->dims[l + 1];
        if (in < 1 || out < 1 || !h->kernel[l] || !h->bias[l]) { rc = dfail(SFX_ERR_ARG, "bad layer"); break; }
        rc = up(d, h->kernel[l], static_cast<size_t>(in) * out, &d->kernel[l]);
        if (rc) break;
        std::vector<float> sc(out), sh(out);
        const bool bn = l < h->n_layers - 1 && h->bn_gamma && h->bn_gamma[l];
        for (int n = 0; n < out; ++n) {
            if (bn) {       // Keras BatchNormalization inference: gamma * (x - mean) / sqrt(var + eps) + beta, x = z + bias
                const double s = static_cast<double>(h->bn_gamma[l][n]) / std::sqrt(static_cast<double>(h->bn_var[l][n]) + h->bn_eps);
                sc[n] = static_cast<float>(s);
                sh[n] = static_cast<float>((static_cast<double>(h->bias[l][n]) - h->bn_mean[l][n]) * s + h->bn_beta[l][n]);
            } else {
                sc[n] = 1.0f;
                sh[n] = h->bias[l][n];
            }
        }
        rc = up(d, sc.data(), sc.size(), &d->scale[l]);
        if (rc) break;
        rc = up(d, sh.data(), sh.size(), &d->shift[l]);
    }
    if (rc == SFX_OK && h->scaler_mean && h->scaler_scale) {
        std::vector<float> mu(h->dims[0]), inv(h->dims[0]);
        for (int k = 0; k < h->dims[0]; ++k) { mu[k] = static_cast<float>(h->scaler_mean[k]); inv[k] = static_cast<float>(1.0 / h->scaler_scale[k]); }
        rc = up(d, mu.data(), mu.size(), &d->pre_mean);
        if (rc == SFX_OK) rc = up(d, inv.data(), inv.size(), &d->pre_inv);
    }
    if (rc != SFX_OK) { sfx_dnn_destroy(d); return rc; }
    *handle = d;
    return SFX_OK;
}

size_t sfx_dnn_workspace_bytes(void* handle, int32_t B) {
    const Dnn* d = static_cast<const Dnn*>(handle);
    if (!d || B < 0) return 0;
    int w = 1;
    for (int l = 1; l < d->n_layers; ++l) w = d->dims[l] > w ? d->dims[l] : w;
    return 2 * static_cast<size_t>(B) * w * sizeof(float) + 256;
}

int sfx_dnn_forward(void* handle, const float* feats, int64_t feat_stride, int32_t B, float* probs, int64_t probs_stride,
                    float* tap, int64_t tap_stride, void* workspace, size_t workspace_bytes, void* stream) {
    Dnn* d = static_cast<Dnn*>(handle);
    if (!d) return dfail(SFX_ERR_ARG, "null handle");
    if (B < 0) return dfail(SFX_ERR_ARG, "B < 0");
    if (B == 0) return SFX_OK;
    if (!feats || !probs || !workspace) return dfail(SFX_ERR_ARG, "null buffer");
    if (workspace_bytes < sfx_dnn_workspace_bytes(handle, B)) return dfail(SFX_ERR_WORKSPACE, "workspace too small");
    cudaSetDevice(d->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int w = 1;
    for (int l = 1; l < d->n_layers; ++l) w = d->dims[l] > w ? d->dims[l] : w;
    float* buf[2] = {static_cast<float*>(workspace), static_cast<float*>(workspace) + static_cast<size_t>(B) * w};
    const float* in = feats;
    long long ldin = feat_stride;
    for (int l = 0; l < d->n_layers; ++l) {
        const int K = d->dims[l], N = d->dims[l + 1];
        const bool last = l == d->n_layers - 1;
        const bool to_tap = (l == d->n_layers - 2) && tap != nullptr;
        float* out = last ? probs : (to_tap ? tap : buf[l & 1]);
        const long long ldo = last ? probs_stride : (to_tap ? tap_stride : N);
        dim3 grid((N + 63) / 64, (B + 63) / 64);
        const float* pm = l == 0 ? d->pre_mean : nullptr;
        const float* pi = l == 0 ? d->pre_inv : nullptr;
        if (last) dense_affine_kernel<false><<<grid, 256, 0, st>>>(in, ldin, d->kernel[l], d->scale[l], d->shift[l], pm, pi, out, ldo, B, N, K);
        else      dense_affine_kernel<true><<<grid, 256, 0, st>>>(in, ldin, d->kernel[l], d->scale[l], d->shift[l], pm, pi, out, ldo, B, N, K);
        in = out;
        ldin = ldo;
    }
    softmax_rows_kernel<<<(B + 7) / 8, 256, 0, st>>>(probs, probs_stride, B, d->dims[d->n_layers]);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return dfail(SFX_ERR_CUDA, std::string("dnn launch: ") + cudaGetErrorString(e));
    return SFX_OK;
}

int sfx_dnn_launches_per_forward(void* handle) {
    const Dnn* d = static_cast<const Dnn*>(handle);
    return d ? d->n_layers + 1 : 0;
}

}  // extern "C"
