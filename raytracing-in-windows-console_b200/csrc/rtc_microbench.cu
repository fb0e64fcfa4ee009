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

// variant 2: FFMA2 whose multiplicand is a scalar broadcast (the `.F32` operand form ptxas picks
// when both halves of a packed operand are the same register) -- the form the ray kernel uses.
__global__ void __launch_bounds__(kPeakThreads, 1)
ffma2_bcast_peak_kernel(int iters, float seed, float* __restrict__ sink)
{
    f32x2 a[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) a[i] = pack2(seed + (float)(threadIdx.x + i), seed - (float)i);
    const float b = 0.999f + seed;
    const f32x2 c = pack2(1.0e-3f + seed, 2.0e-3f + seed);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < kInner; ++k) {
#pragma unroll
            for (int i = 0; i < kChains; ++i) a[i] = fma2(pack2(b, b), a[i], c);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kChains; ++i) { float lo, hi; unpack2(a[i], lo, hi); s += lo + hi; }
    if (s == 12345.678f) sink[0] = s;
}

// variant 3: the ray kernel's inner-loop instruction mix with operands in registers (no LDS, no
// branch): per ray FMUL2 + 3 FFMA2 + FMNMX3, 8 rays, "sphere pair" operands rotate through 4 sets.
template <bool ORDERED>
__global__ void __launch_bounds__(kPeakThreads, 1)
raymix_peak_kernel(int iters, float seed, float* __restrict__ sink)
{
    float dx[8], dy[8], dz[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) { dx[r] = seed + 0.01f * (threadIdx.x + r); dy[r] = seed + 0.02f * r; dz[r] = seed + 0.5f; }
    f32x2 ox[4], oy[4], oz[4], nc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { ox[j] = pack2(seed + j, seed - j); oy[j] = pack2(seed + 2 * j, seed + 1.f); oz[j] = pack2(seed - 3.f, seed + j); nc[j] = pack2(-1.f - seed, -2.f - j); }
    float m = -1.0f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                f32x2 t[8];
                if (ORDERED) {          // operand-major order: 8 consecutive packed ops share one sphere operand
#pragma unroll
                    for (int r = 0; r < 8; ++r) t[r] = mul2(pack2(dx[r], dx[r]), ox[j]);
#pragma unroll
                    for (int r = 0; r < 8; ++r) t[r] = fma2(pack2(dy[r], dy[r]), oy[j], t[r]);
#pragma unroll
                    for (int r = 0; r < 8; ++r) t[r] = fma2(pack2(dz[r], dz[r]), oz[j], t[r]);
#pragma unroll
                    for (int r = 0; r < 8; ++r) t[r] = fma2(t[r], t[r], nc[j]);
                } else {
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        f32x2 u = mul2(pack2(dx[r], dx[r]), ox[j]);
                        u = fma2(pack2(dy[r], dy[r]), oy[j], u);
                        u = fma2(pack2(dz[r], dz[r]), oz[j], u);
                        t[r] = fma2(u, u, nc[j]);
                    }
                }
#pragma unroll
                for (int r = 0; r < 8; ++r) { float lo, hi; unpack2(t[r], lo, hi); m = max3(m, lo, hi); }
            }
            // keep the operands loop-variant so nothing is hoisted or merged across halves
#pragma unroll
            for (int j = 0; j < 4; ++j) ox[j] = fma2(ox[j], pack2(1.0000001f, 1.0000001f), pack2(m * 1e-30f, 0.f));
        }
    }
    if (m == 12345.678f) sink[0] = m;
}

// FLOPs executed by one launch of the given variant.
double fp32_peak_flops(int variant, int n_ctas, int iters)
{
    if (variant == 3 || variant == 4)   // 8 pairs x 8 rays x 2 tests x 7 algorithmic FLOP per test (the roofline's own accounting)
        return (double)n_ctas * kPeakThreads * (double)iters * 8.0 * 8.0 * 2.0 * 7.0;
    const double fmas = (double)n_ctas * kPeakThreads * (double)iters * kInner * kChains * (variant >= 1 ? 2.0 : 1.0);
    return 2.0 * fmas;
}

cudaError_t launch_fp32_peak(cudaStream_t st, int variant, int n_ctas, int iters, float* sink)
{
    if (variant == 4) raymix_peak_kernel<true><<<n_ctas, kPeakThreads, 0, st>>>(iters, 0.0f, sink);
    else if (variant == 3) raymix_peak_kernel<false><<<n_ctas, kPeakThreads, 0, st>>>(iters, 0.0f, sink);
    else if (variant == 2) ffma2_bcast_peak_kernel<<<n_ctas, kPeakThreads, 0, st>>>(iters, 0.0f, sink);
    else if (variant == 1) ffma2_peak_kernel<<<n_ctas, kPeakThreads, 0, st>>>(iters, 0.0f, sink);
    else ffma_peak_kernel<<<n_ctas, kPeakThreads, 0, st>>>(iters, 0.0f, sink);
    return cudaGetLastError();
}

}  // namespace rtc
