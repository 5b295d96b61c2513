/* sfx_bench.h -- measurement helpers of libsfx_bench.so (bench.py / tests only; NOT part of the product library
 * libsfx_b200.so and not on any extraction path). */
#ifndef SFX_BENCH_H
#define SFX_BENCH_H

#ifdef __cplusplus
extern "C" {
#endif

/* SURVEY.md 8(d) "measure an FP32 FMA micro-benchmark peak on the box as the compute denominator": FP32 FMA throughput of
 * `device` in TFLOP/s from a register-only FFMA kernel (64 warps/SM), timed with CUDA events on a private stream, best of
 * three launches.  Returns 0 on success, a negative sfx_status otherwise. */
int sfx_measure_fp32_peak(int device, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* SFX_BENCH_H */
