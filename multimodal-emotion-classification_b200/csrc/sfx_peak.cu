// sfx_peak.cu -- measurement helper ("measure an FP32 FMA micro-benchmark peak on the box as the compute
// denominator").  The extractor is bound by FP32 issue slots and shared memory, not by HBM, so bench.py reports the
// algorithmic flop rate against this number next to the HBM roofline BASELINE.json asks for.  Built into its own library
// (libsfx_bench.so, include/sfx_bench.h): not part of libsfx_b200.so.
#include <cuda_runtime.h>
#include <string>

#include "../../include/sfx.h"
#include "../../include/sfx_bench.h"

namespace {

constexpr int kChains = 32, kIters = 16384, kPeakThreads = 256, kCtasPerSm = 8;

// 32 independent FFMA chains per thread, 64 warps per SM: the FP32 pipes are the only busy resource
__global__ void __launch_bounds__(kPeakThreads) fp32_fma_peak_kernel(float* __restrict__ out, const float mul, const float add) {
    float a[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) a[i] = mul + static_cast<float>(i) + 1e-6f * threadIdx.x;
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i) a[i] = fmaf(a[i], mul, add);
    }
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += a[i];
    out[static_cast<size_t>(blockIdx.x) * kPeakThreads + threadIdx.x] = s;
}

}  // namespace

extern "C" int sfx_measure_fp32_peak(int device, double* tflops) {
    if (!tflops) return SFX_ERR_ARG;
    *tflops = 0.0;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); return SFX_ERR_CUDA; }
    if (cudaSetDevice(device) != cudaSuccess) return SFX_ERR_CUDA;
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return SFX_ERR_CUDA;
    const int grid = sms * kCtasPerSm;
    float* d = nullptr;
    cudaStream_t st = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int rc = SFX_ERR_CUDA;
    float best_ms = 0.0f;
    if (cudaMalloc(&d, sizeof(float) * static_cast<size_t>(grid) * kPeakThreads) != cudaSuccess) goto done;
    if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) goto done;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) goto done;
    for (int rep = 0; rep < 4; ++rep) {                       // first launch = warm-up, then the best of three
        if (cudaEventRecord(e0, st) != cudaSuccess) goto done;
        fp32_fma_peak_kernel<<<grid, kPeakThreads, 0, st>>>(d, 0.99993896484375f, 0.25f);
        if (cudaEventRecord(e1, st) != cudaSuccess || cudaEventSynchronize(e1) != cudaSuccess) goto done;
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, e0, e1) != cudaSuccess) goto done;
        if (rep > 0 && (best_ms == 0.0f || ms < best_ms)) best_ms = ms;
    }
    if (best_ms > 0.0f) {
        const double flops = 2.0 * kChains * kIters * static_cast<double>(kPeakThreads) * grid;   // FMA = 2 flops
        *tflops = flops / (best_ms * 1e-3) * 1e-12;
        rc = SFX_OK;
    }
done:
    if (rc != SFX_OK) cudaGetLastError();
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (st) cudaStreamDestroy(st);
    if (d) cudaFree(d);
    return rc;
}
