// Micro-benchmark behind the ingest design: 64 pinned H2D copies of ~276 KB (the CIGAR op arrays of a 10 Mbp contig as
// 16 parse segments x 4 arrays) against one copy of the same bytes, with the pinned buffers allocated and written
// either by the main thread or by short-lived worker threads (what the parallel record parse does).
//   nvcc -O2 -o /tmp/h2d h2d_small_copies.cu && /tmp/h2d
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("err %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
int main() {
    cudaStream_t s; CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    const size_t n = 69000 * 4; // bytes per copy
    const int C = 64;
    for (int variant = 0; variant < 4; variant++) {
        std::vector<void*> h(C);
        const bool threaded = variant & 1, rewrite = variant & 2;
        size_t cap = 131072 * 4;
        auto mk = [&](int i) { CK(cudaHostAlloc(&h[i], cap, cudaHostAllocPortable)); memset(h[i], 1, cap); };
        if (threaded) { std::vector<std::thread> th; for (int i = 0; i < C; i++) th.emplace_back(mk, i); for (auto &t : th) t.join(); }
        else for (int i = 0; i < C; i++) mk(i);
        char *d; CK(cudaMalloc(&d, n * C));
        cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
        for (int rep = 0; rep < 3; rep++) {
            if (rewrite) {  // fresh data written by worker threads right before the copies (dirty lines in other cores' caches)
                std::vector<std::thread> th;
                for (int i = 0; i < C; i++) th.emplace_back([&, i]() { uint32_t *p = (uint32_t *)h[i]; for (size_t x = 0; x < n / 4; x++) p[x] = (uint32_t)(x * 2654435761u + rep); });
                for (auto &t : th) t.join();
            }
            CK(cudaEventRecord(a, s));
            for (int i = 0; i < C; i++) CK(cudaMemcpyAsync(d + i * n, h[i], n, cudaMemcpyHostToDevice, s));
            CK(cudaEventRecord(b, s));
            CK(cudaStreamSynchronize(s));
            float ms; CK(cudaEventElapsedTime(&ms, a, b));
            printf("alloc by %s, %s: rep %d: %d copies x %zu B: %.3f ms (%.1f GB/s)\n", threaded ? "worker threads" : "main thread", rewrite ? "rewritten by threads before each rep" : "static data", rep, C, n, ms, C * n / ms / 1e6);
        }
        for (auto p : h) cudaFreeHost(p);
        cudaFree(d);
    }
    return 0;
}
