// Micro-benchmark: issue rate of packed FP32 (fma.rn.f32x2 / add.f32x2) vs scalar FFMA/FADD on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu && ./f32x2
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pk(float a, float b) {
    return (static_cast<unsigned long long>(__float_as_uint(b)) << 32) | __float_as_uint(a);
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

constexpr int kIters = 4096, kChains = 16;

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float seed) {
    const float t = seed + threadIdx.x * 1e-6f;
    if (MODE == 0) {                       // scalar FFMA, 2*kChains independent chains, register twiddle
        float a[2 * kChains];
#pragma unroll
        for (int i = 0; i < 2 * kChains; ++i) a[i] = t + i;
        for (int it = 0; it < kIters; ++it) {
#pragma unroll
            for (int i = 0; i < 2 * kChains; ++i) a[i] = fmaf(a[i], t, 0.25f);
        }
        float s = 0;
#pragma unroll
        for (int i = 0; i < 2 * kChains; ++i) s += a[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else if (MODE == 1) {                // packed FFMA2, kChains chains of 2 lanes (same flops as MODE 0)
        unsigned long long a[kChains];
        const unsigned long long tt = pk(t, t), cc = pk(0.25f, 0.25f);
#pragma unroll
        for (int i = 0; i < kChains; ++i) a[i] = pk(t + i, t - i);
        for (int it = 0; it < kIters; ++it) {
#pragma unroll
            for (int i = 0; i < kChains; ++i) a[i] = fma2(a[i], tt, cc);
        }
        unsigned long long s = 0;
#pragma unroll
        for (int i = 0; i < kChains; ++i) s ^= a[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(static_cast<unsigned>(s ^ (s >> 32)));
    } else if (MODE == 2) {                // scalar FADD
        float a[2 * kChains];
#pragma unroll
        for (int i = 0; i < 2 * kChains; ++i) a[i] = t + i;
        for (int it = 0; it < kIters; ++it) {
#pragma unroll
            for (int i = 0; i < 2 * kChains; ++i) a[i] = a[i] + t;
        }
        float s = 0;
#pragma unroll
        for (int i = 0; i < 2 * kChains; ++i) s += a[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else if (MODE == 3) {                // packed FADD2
        unsigned long long a[kChains];
        const unsigned long long tt = pk(t, t);
#pragma unroll
        for (int i = 0; i < kChains; ++i) a[i] = pk(t + i, t - i);
        for (int it = 0; it < kIters; ++it) {
#pragma unroll
            for (int i = 0; i < kChains; ++i) a[i] = add2(a[i], tt);
        }
        unsigned long long s = 0;
#pragma unroll
        for (int i = 0; i < kChains; ++i) s ^= a[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(static_cast<unsigned>(s ^ (s >> 32)));
    } else if (MODE == 4) {                // scalar FFMA with immediate multiplier (imm-form)
        float a[2 * kChains];
#pragma unroll
        for (int i = 0; i < 2 * kChains; ++i) a[i] = t + i;
        for (int it = 0; it < kIters; ++it) {
#pragma unroll
            for (int i = 0; i < 2 * kChains; ++i) a[i] = fmaf(a[i], 0.999f, 0.25f);
        }
        float s = 0;
#pragma unroll
        for (int i = 0; i < 2 * kChains; ++i) s += a[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else if (MODE == 5) {                // packed FFMA2 with compile-time constant operands
        unsigned long long a[kChains];
#pragma unroll
        for (int i = 0; i < kChains; ++i) a[i] = pk(t + i, t - i);
        for (int it = 0; it < kIters; ++it) {
#pragma unroll
            for (int i = 0; i < kChains; ++i) {
                unsigned long long d;
                asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a[i]), "l"(0x3f7fbe773f7fbe77ull), "l"(0x3e8000003e800000ull));
                a[i] = d;
            }
        }
        unsigned long long s = 0;
#pragma unroll
        for (int i = 0; i < kChains; ++i) s ^= a[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(static_cast<unsigned>(s ^ (s >> 32)));
    }
}

template <int MODE>
float run(float* d, int grid) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, 256>>>(d, 0.5f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<grid, 256>>>(d, 0.5f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = sms * 8;           // 8 CTAs x 8 warps = 64 warps/SM
    float* d;
    cudaMalloc(&d, sizeof(float) * grid * 256);
    const double flops = 2.0 * 2 * kChains * kIters * 256.0 * grid;     // FMA = 2 flops
    const char* names[] = {"FFMA  reg", "FFMA2 reg", "FADD  reg", "FADD2 reg", "FFMA  imm", "FFMA2 const"};
    float ms[6] = {run<0>(d, grid), run<1>(d, grid), run<2>(d, grid), run<3>(d, grid), run<4>(d, grid), run<5>(d, grid)};
    for (int i = 0; i < 6; ++i) {
        const double ops = flops / ((i == 2 || i == 3) ? 2 : 1);
        printf("%-12s %8.3f ms  %8.2f T%s/s  lane-ops/clk/SM @1.965GHz = %.1f\n", names[i], ms[i], ops / ms[i] * 1e-9,
               (i == 2 || i == 3) ? "add" : "flop", ops / ((i == 2 || i == 3) ? 1 : 2) / (ms[i] * 1e-3) / 1.965e9 / sms);
    }
    if (cudaGetLastError() != cudaSuccess) { printf("CUDA error\n"); return 1; }
    return 0;
}
