// rtc_microbench.cu -- FP32-pipe peak probes: the measured denominator of the ray kernel's
// roofline (BASELINE.md "plus an on-box FFMA microbenchmark").
//   variant 0: scalar FFMA, 16 independent chains per thread.
//   variant 1: packed FFMA2 (fma.rn.f32x2), 16 independent chains per thread (32 FMAs).
// Each CTA has 512 threads; n_ctas = a multiple of the SM count.  FLOPs = 2 per FMA.
#include "rtc_device.cuh"
#include "rtc_kernels.h"

namespace rtc {

constexpr int kPeakThreads = 512;
constexpr int kChains = 16;
constexpr int kInner = 64;

__global__ void __launch_bounds__(kPeakThreads, 1)
ffma_peak_kernel(int iters, float seed, float* __restrict__ sink)
{
    float a[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) a[i] = seed + (float)(threadIdx.x + i);
    const float b = 0.999f + seed, c = 1.0e-3f + seed;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < kInner; ++k) {
#pragma unroll
            for (int i = 0; i < kChains; ++i) a[i] = fmaf(a[i], b, c);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += a[i];
    if (s == 12345.678f) sink[0] = s;
}

__global__ void __launch_bounds__(kPeakThreads, 1)
ffma2_peak_kernel(int iters, float seed, float* __restrict__ sink)
{
    f32x2 a[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) a[i] = pack2(seed + (float)(threadIdx.x + i), seed - (float)i);
    const f32x2 b = pack2(0.999f + seed, 0.998f + seed), c = pack2(1.0e-3f + seed, 2.0e-3f + seed);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < kInner; ++k) {
#pragma unroll
            for (int i = 0; i < kChains; ++i) a[i] = fma2(a[i], b, c);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kChains; ++i) { float lo, hi; unpack2(a[i], lo, hi); s += lo + hi; }
    if (s == 12345.678f) sink[0] = s;
}

// FLOPs executed by one launch of the given variant.
double fp32_peak_flops(int variant, int n_ctas, int iters)
{
    const double fmas = (double)n_ctas * kPeakThreads * (double)iters * kInner * kChains * (variant == 1 ? 2.0 : 1.0);
    return 2.0 * fmas;
}

cudaError_t launch_fp32_peak(cudaStream_t st, int variant, int n_ctas, int iters, float* sink)
{
    if (variant == 1) ffma2_peak_kernel<<<n_ctas, kPeakThreads, 0, st>>>(iters, 0.0f, sink);
    else ffma_peak_kernel<<<n_ctas, kPeakThreads, 0, st>>>(iters, 0.0f, sink);
    return cudaGetLastError();
}

}  // namespace rtc
