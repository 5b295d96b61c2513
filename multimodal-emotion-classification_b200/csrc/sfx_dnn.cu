// sfx_dnn.cu -- device-resident scaler + speech DNN forward (scope row f1): the consumer of the 56-d features in
// the reference's inference/speech_inference.py:66-76,85-105 (architecture: model_training/train_speech_model.py:55-90).
// StandardScaler -> [Dense + BatchNorm(eps) + ReLU] x (L-1) -> Dense + softmax, with the last hidden activation (Keras
// layers[-3]) returned as the fusion feature tap.
//
// One launch (dnn_fused_kernel): a CTA carries 32 feature rows through every layer with the activations in shared memory;
// the dense layers run on the tensor cores (mma.sync m16n8k16, FP32 accumulate) at float32-equivalent accuracy: weights and
// activations are split x = hi + 2^-11 lo into two FP16 values (22 significant bits) and a layer is hi.hi + 2^-11 (hi.lo +
// lo.hi); the lo.lo term (2^-22 relative) is dropped.  Stated tolerance against the float32 restatement
// (oracle/speech_dnn.py, tests/test_dnn.py): probabilities 2e-5 absolute, 64-d tap 5e-4 relative to its maximum -- the
// bounds the previous all-FP32 SIMT kernels were tested with.  Weights are pre-arranged at create time as ready-made B
// fragments (one coalesced 64-bit load per lane, MMA and K step), A fragments come from shared memory with ldmatrix.
// Models wider than 512 fall back to the per-layer FP32 kernels below (a 64x64x16 shared-memory tiled SIMT GEMM with the
// per-channel affine + ReLU fused in).
#include <cuda_runtime.h>
#include <cuda_fp16.h>
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
    uint2* bhi[kMaxLayers] = {nullptr};      // fused path: B fragments of the kernel, hi / 2^11 * lo halves
    uint2* blo[kMaxLayers] = {nullptr};
    bool fused = false;
    std::vector<void*> allocs;
};

// ---------------------------------------------------------------------------------------------- fused tensor-core forward
constexpr int kFR = 32;                      // feature rows per CTA (two 16-row MMA tiles)
constexpr int kFMaxW = 512;                  // widest layer the fused kernel takes
constexpr int kFLd = kFMaxW + 8;             // halves per activation row in shared memory (1040 B: conflict-free ldmatrix)
static_assert(kFMaxW % 64 == 0, "layer widths are padded to 64 inside the activation rows");
constexpr int kFThreads = 256;

struct FusedLayer { const uint2* bhi; const uint2* blo; const float* scale; const float* shift; int K, N, Kp, Np; };
struct FusedParams {
    FusedLayer L[kMaxLayers];
    int n_layers;
    const float* pre_mean; const float* pre_inv;
    const float* feats; long long feat_stride;
    int B;
    float* probs; long long probs_stride;
    float* tap; long long tap_stride;
};

__host__ __device__ inline int round16(int x) { return (x + 15) & ~15; }
__host__ __device__ inline int round64(int x) { return (x + 63) & ~63; }     // K is padded to 4 MMA K steps: the weight loop prefetches four steps

__device__ __forceinline__ void ldmatrix_x4(unsigned (&r)[4], const __half* ptr) {
    const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(ptr));
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// x = hi + 2^-11 * lo with hi, lo FP16
__device__ __forceinline__ void split_h(float v, __half& hi, __half& lo) {
    hi = __float2half_rn(v);
    lo = __float2half_rn((v - __half2float(hi)) * 2048.0f);
}

__global__ void __launch_bounds__(kFThreads) dnn_fused_kernel(const __grid_constant__ FusedParams p) {
    extern __shared__ __align__(16) unsigned char sm[];
    __half* act = reinterpret_cast<__half*>(sm);                       // [2 buffers][2 hi/lo][kFR][kFLd]
    float* logits = reinterpret_cast<float*>(act + 4 * kFR * kFLd);    // [kFR][32]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int row0 = blockIdx.x * kFR;
    auto plane = [&](int buf, int hl) { return act + (buf * 2 + hl) * kFR * kFLd; };

    // scaled features -> buffer 0 (StandardScaler.transform, speech_inference.py:66-67), zero padded to a multiple of 16
    {
        const int K = p.L[0].K, Kp = p.L[0].Kp;
        __half* hi = plane(0, 0);
        __half* lo = plane(0, 1);
        for (int i = tid; i < kFR * Kp; i += kFThreads) {
            const int r = i / Kp, k = i - r * Kp;
            float v = 0.0f;
            if (row0 + r < p.B && k < K) {
                v = p.feats[static_cast<long long>(row0 + r) * p.feat_stride + k];
                if (p.pre_mean) v = (v - p.pre_mean[k]) * p.pre_inv[k];
            }
            split_h(v, hi[r * kFLd + k], lo[r * kFLd + k]);
        }
    }
    __syncthreads();

    for (int l = 0; l < p.n_layers; ++l) {
        const FusedLayer& L = p.L[l];
        const bool last = l == p.n_layers - 1;
        const bool to_tap = l == p.n_layers - 2 && p.tap != nullptr;
        const __half* in_hi = plane(l & 1, 0);
        const __half* in_lo = plane(l & 1, 1);
        __half* out_hi = plane((l + 1) & 1, 0);
        __half* out_lo = plane((l + 1) & 1, 1);
        const int ntiles = L.Np >> 3, ksteps = L.Kp >> 4;
        // this lane's ldmatrix row / column offset inside a 16 x 16 A tile
        const int a_off = (lane & 15) * kFLd + (lane >> 4) * 8;
        for (int j = warp; j < ntiles; j += kFThreads / 32) {
            float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};      // hi . hi
            float acx[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};      // 2^11 * (hi . lo + lo . hi)
            const uint2* bh = L.bhi + static_cast<size_t>(j) * ksteps * 32 + lane;
            const uint2* bl = L.blo + static_cast<size_t>(j) * ksteps * 32 + lane;
            // weights four K steps ahead of the MMAs that use them (one L2 round trip per 24 MMAs instead of per 6)
            uint2 wh[4], wl[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { wh[q] = __ldg(bh + q * 32); wl[q] = __ldg(bl + q * 32); }
            for (int s0 = 0; s0 < ksteps; s0 += 4) {
                uint2 nwh[4], nwl[4];
                const int sn = min(s0 + 4, ksteps - 4);
#pragma unroll
                for (int q = 0; q < 4; ++q) { nwh[q] = __ldg(bh + (sn + q) * 32); nwl[q] = __ldg(bl + (sn + q) * 32); }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
                        unsigned ah[4], al[4];
                        ldmatrix_x4(ah, in_hi + m * 16 * kFLd + (s0 + q) * 16 + a_off);
                        ldmatrix_x4(al, in_lo + m * 16 * kFLd + (s0 + q) * 16 + a_off);
                        mma16816(acc[m], ah, wh[q].x, wh[q].y);
                        mma16816(acx[m], ah, wl[q].x, wl[q].y);
                        mma16816(acx[m], al, wh[q].x, wh[q].y);
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) { wh[q] = nwh[q]; wl[q] = nwl[q]; }
            }
            // epilogue: D fragment = rows g, g+8 of the 16-row tile, columns 8j + 2*t4, +1
            const int c0 = j * 8 + 2 * t4;
            float sc[2], sh[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const bool in = c0 + e < L.N;
                sc[e] = in ? L.scale[c0 + e] : 0.0f;
                sh[e] = in ? L.shift[c0 + e] : 0.0f;
            }
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int r = m * 16 + h * 8 + g;
                    float v[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        v[e] = fmaf(fmaf(acx[m][2 * h + e], 1.0f / 2048.0f, acc[m][2 * h + e]), sc[e], sh[e]);
                        if (!last) v[e] = fmaxf(v[e], 0.0f);
                    }
                    if (last) {
                        if (c0 < 32) { logits[r * 32 + c0] = v[0]; logits[r * 32 + c0 + 1] = v[1]; }
                    } else {
                        __half h0, l0, h1, l1;
                        split_h(v[0], h0, l0);
                        split_h(v[1], h1, l1);
                        *reinterpret_cast<__half2*>(out_hi + r * kFLd + c0) = __halves2half2(h0, h1);
                        *reinterpret_cast<__half2*>(out_lo + r * kFLd + c0) = __halves2half2(l0, l1);
                        if (to_tap && row0 + r < p.B) {
                            float* t = p.tap + static_cast<long long>(row0 + r) * p.tap_stride;
                            if (c0 < L.N) t[c0] = v[0];
                            if (c0 + 1 < L.N) t[c0 + 1] = v[1];
                        }
                    }
                }
        }
        __syncthreads();
    }
    // softmax of the last layer's logits (one thread per row: 7 classes)
    if (tid < kFR && row0 + tid < p.B) {
        const int N = p.L[p.n_layers - 1].N;
        const float* z = logits + tid * 32;
        float mx = z[0];
        for (int c = 1; c < N; ++c) mx = fmaxf(mx, z[c]);
        float e[32], sum = 0.0f;
        for (int c = 0; c < N; ++c) { e[c] = expf(z[c] - mx); sum += e[c]; }
        float* o = p.probs + static_cast<long long>(row0 + tid) * p.probs_stride;
        for (int c = 0; c < N; ++c) o[c] = e[c] / sum;
    }
}

size_t fused_smem() { return sizeof(__half) * 4 * kFR * kFLd + sizeof(float) * kFR * 32; }

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
    // fused tensor-core path: B fragments of every kernel (hi / lo halves), [n tile][k step][lane] -> (b0, b1)
    bool fits = true;
    for (int i = 0; i <= h->n_layers; ++i) fits &= h->dims[i] <= kFMaxW;
    if (rc == SFX_OK && fits) {
        for (int l = 0; l < h->n_layers && rc == SFX_OK; ++l) {
            const int K = h->dims[l], N = h->dims[l + 1], Kp = round64(K), Np = l + 1 < h->n_layers ? round64(N) : round16(N);
            const int ntiles = Np / 8, ksteps = Kp / 16;
            std::vector<uint2> fh(static_cast<size_t>(ntiles) * ksteps * 32), fl(fh.size());
            auto split = [&](int k, int n, unsigned short& hi, unsigned short& lo) {
                const float w = (k < K && n < N) ? h->kernel[l][static_cast<size_t>(k) * N + n] : 0.0f;
                const __half a = __float2half_rn(w);
                const __half b = __float2half_rn((w - __half2float(a)) * 2048.0f);
                hi = __half_as_ushort(a);
                lo = __half_as_ushort(b);
            };
            for (int j = 0; j < ntiles; ++j)
                for (int sidx = 0; sidx < ksteps; ++sidx)
                    for (int ln = 0; ln < 32; ++ln) {
                        const int gg = ln >> 2, tt = ln & 3, n = j * 8 + gg, k0 = sidx * 16 + 2 * tt;
                        unsigned short h00, l00, h01, l01, h10, l10, h11, l11;
                        split(k0, n, h00, l00); split(k0 + 1, n, h01, l01);
                        split(k0 + 8, n, h10, l10); split(k0 + 9, n, h11, l11);
                        const size_t at = (static_cast<size_t>(j) * ksteps + sidx) * 32 + ln;
                        fh[at] = make_uint2(h00 | (static_cast<unsigned>(h01) << 16), h10 | (static_cast<unsigned>(h11) << 16));
                        fl[at] = make_uint2(l00 | (static_cast<unsigned>(l01) << 16), l10 | (static_cast<unsigned>(l11) << 16));
                    }
            rc = up(d, fh.data(), fh.size(), &d->bhi[l]);
            if (rc == SFX_OK) rc = up(d, fl.data(), fl.size(), &d->blo[l]);
        }
        if (rc == SFX_OK) {
            cudaError_t e2 = cudaFuncSetAttribute(dnn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(fused_smem()));
            if (e2 == cudaSuccess) d->fused = true;
            else cudaGetLastError();
        }
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
    if (d->fused) {
        FusedParams fp{};
        fp.n_layers = d->n_layers;
        for (int l = 0; l < d->n_layers; ++l)
            fp.L[l] = FusedLayer{d->bhi[l], d->blo[l], d->scale[l], d->shift[l], d->dims[l], d->dims[l + 1], round64(d->dims[l]),
                                 l + 1 < d->n_layers ? round64(d->dims[l + 1]) : round16(d->dims[l + 1])};
        fp.pre_mean = d->pre_mean; fp.pre_inv = d->pre_inv;
        fp.feats = feats; fp.feat_stride = feat_stride; fp.B = B;
        fp.probs = probs; fp.probs_stride = probs_stride; fp.tap = tap; fp.tap_stride = tap_stride;
        dnn_fused_kernel<<<(B + kFR - 1) / kFR, kFThreads, fused_smem(), st>>>(fp);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return dfail(SFX_ERR_CUDA, std::string("dnn launch: ") + cudaGetErrorString(e));
        return SFX_OK;
    }
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
    return d ? (d->fused ? 1 : d->n_layers + 1) : 0;
}

}  // extern "C"
