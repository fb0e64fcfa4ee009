// probe_d2h_fanin.cu -- how fast can N GPUs write into ONE pinned host buffer at the same time?
// (The multi-GPU frame driver's end-to-end ceiling: eight band streams of ~2.6 MB land in one host frame per 4K frame.)
//   nvcc -O2 -o probe_d2h_fanin probe_d2h_fanin.cu && ./probe_d2h_fanin
// For N = 1, 2, 4, 8 devices, chunk = 2.6 MB and 32 MB, host memory = cudaHostAlloc / registered anonymous memory with
// transparent huge pages: N threads, each copying its chunk `iters` times into its own region; aggregate GB/s.
#include <cuda_runtime.h>
#include <sys/mman.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

static double run(int n, size_t chunk, char* host, int iters)
{
    std::vector<char*> dev(n);
    std::vector<cudaStream_t> st(n);
    for (int g = 0; g < n; ++g) {
        cudaSetDevice(g);
        cudaMalloc(&dev[g], chunk);
        cudaMemset(dev[g], g + 1, chunk);
        cudaStreamCreateWithFlags(&st[g], cudaStreamNonBlocking);
        cudaDeviceSynchronize();
    }
    auto body = [&](int g, int it) {
        cudaSetDevice(g);
        for (int i = 0; i < it; ++i) cudaMemcpyAsync(host + (size_t)g * chunk, dev[g], chunk, cudaMemcpyDeviceToHost, st[g]);
        cudaStreamSynchronize(st[g]);
    };
    { std::vector<std::thread> th; for (int g = 0; g < n; ++g) th.emplace_back(body, g, 3); for (auto& t : th) t.join(); }
    const auto t0 = std::chrono::steady_clock::now();
    { std::vector<std::thread> th; for (int g = 0; g < n; ++g) th.emplace_back(body, g, iters); for (auto& t : th) t.join(); }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (int g = 0; g < n; ++g) { cudaSetDevice(g); cudaFree(dev[g]); cudaStreamDestroy(st[g]); }
    return (double)n * chunk * iters / secs / 1e9;
}

int main()
{
    int n_dev = 0;
    cudaGetDeviceCount(&n_dev);
    const size_t chunks[2] = {2600000, 32u << 20};
    const size_t total = 8 * chunks[1];
    char* pinned = nullptr;
    cudaHostAlloc((void**)&pinned, total, cudaHostAllocPortable);
    char* thp = (char*)mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    madvise(thp, total, MADV_HUGEPAGE);
    memset(thp, 0, total);
    const bool thp_ok = cudaHostRegister(thp, total, cudaHostRegisterPortable) == cudaSuccess;
    printf("| devices | chunk | cudaHostAlloc GB/s | registered THP GB/s |\n|---|---|---|---|\n");
    for (int n = 1; n <= n_dev && n <= 8; n *= 2)
        for (size_t c : chunks) {
            const int iters = c > (4u << 20) ? 40 : 400;
            const double a = run(n, c, pinned, iters);
            const double b = thp_ok ? run(n, c, thp, iters) : 0.0;
            printf("| %d | %.1f MB | %.1f | %.1f |\n", n, c / 1e6, a, b);
        }
    return 0;
}
