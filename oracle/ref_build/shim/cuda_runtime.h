// Build shim (test infrastructure, CPU build only): a fake <cuda_runtime.h> that lets
// the reference's .cu files compile as ordinary C++.  Kernels become plain functions;
// `kernel CUDA_KERNEL(g,b)(args)` is rewritten (see wrap_all.cpp) into
// `kernel * fakecuda::Launch{g,b}(args)`, which loops grid x block on the host and
// sets thread-local blockIdx/threadIdx the way the hardware would.
#pragma once
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <tuple>
#include <thread>
#include <vector>
#include <utility>

#define __host__
#define __device__
#define __global__
#define __constant__ static

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct fake_uint3 { unsigned x = 0, y = 0, z = 0; };
extern thread_local fake_uint3 blockIdx, threadIdx;
extern thread_local dim3 blockDim, gridDim;

typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
inline const char* cudaGetErrorString(cudaError_t) { return "fake-cuda error"; }
template <class T> inline cudaError_t cudaMalloc(T** p, size_t n) { *p = (T*)std::malloc(n); return *p ? 0 : 2; }
inline cudaError_t cudaFree(void* p) { std::free(p); return 0; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return 0; }
inline cudaError_t cudaDeviceSynchronize() { return 0; }

namespace fakecuda {
extern int g_threads;   // host threads used to run a "grid" (1 = serial)
template <class... A> struct Bound { dim3 g, b; std::tuple<A...> a; };
struct Launch {
    dim3 g, b;
    template <class... A> Bound<A...> operator()(A... a) const { return Bound<A...>{g, b, std::tuple<A...>(a...)}; }
};
template <class... P, class... A>
void operator*(void (*k)(P...), Bound<A...> bd) {
    const unsigned long long bt = (unsigned long long)bd.b.x * bd.b.y * bd.b.z;
    // CUDA rejects the launch (cudaErrorInvalidConfiguration) and the kernel never runs.
    if (bt == 0 || bt > 1024 || bd.g.x == 0 || bd.g.y == 0) return;
    const unsigned long long nblocks = (unsigned long long)bd.g.x * bd.g.y;
    auto run = [&](unsigned long long lo, unsigned long long hi) {
        gridDim = bd.g; blockDim = bd.b;
        for (unsigned long long blk = lo; blk < hi; ++blk) {
            blockIdx.x = (unsigned)(blk % bd.g.x); blockIdx.y = (unsigned)(blk / bd.g.x); blockIdx.z = 0;
            for (unsigned ty = 0; ty < bd.b.y; ++ty)
                for (unsigned tx = 0; tx < bd.b.x; ++tx) {
                    threadIdx.x = tx; threadIdx.y = ty; threadIdx.z = 0;
                    std::apply(k, bd.a);
                }
        }
    };
    int T = g_threads < 1 ? 1 : g_threads;
    if ((unsigned long long)T > nblocks) T = (int)nblocks;
    if (T <= 1) { run(0, nblocks); return; }
    std::vector<std::thread> pool;
    for (int t = 0; t < T; ++t)
        pool.emplace_back(run, nblocks * t / T, nblocks * (t + 1) / T);
    for (auto& th : pool) th.join();
}
}  // namespace fakecuda
using fakecuda::operator*;
