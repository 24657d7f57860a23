// How many DRAM bytes does one random 32-byte load cost on B200, and which load qualifier changes it?
// 2^26 random 32-byte (one sector) loads over a 1 GiB buffer, one per thread iteration, for several PTX forms of the load.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_fetch l2_fetch.cu
// run:   ./l2_fetch            (CUDA-event times)
//        ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum ./l2_fetch     (bytes per variant)
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

template <int V>
__device__ __forceinline__ void ld32(const uint64_t *p, uint64_t v[4]) {
    if (V == 0) asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]) : "l"(p));
    if (V == 1) asm volatile("ld.global.nc.L2::evict_first.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]) : "l"(p));
    if (V == 2) asm volatile("ld.global.nc.L2::64B.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]) : "l"(p));
    if (V == 3) asm volatile("ld.global.nc.L2::128B.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]) : "l"(p));
    if (V == 4) asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]) : "l"(p));
    if (V == 5) {  // two 128-bit loads (the round-1 form)
        asm volatile("ld.global.nc.v2.u64 {%0,%1}, [%2];" : "=l"(v[0]), "=l"(v[1]) : "l"(p));
        asm volatile("ld.global.nc.v2.u64 {%0,%1}, [%2];" : "=l"(v[2]), "=l"(v[3]) : "l"(p + 2));
    }
    if (V == 6) asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]) : "l"(p));
    if (V == 7) asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.L2::64B.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]) : "l"(p));
}
template <int V>
__global__ void __launch_bounds__(256) k(const uint64_t *buf, uint64_t n_sectors, uint64_t n_loads, uint64_t *sink) {
    uint64_t acc = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_loads; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t h = (i + 1) * 0x9E3779B97F4A7C15ULL;
        h ^= h >> 29;
        h *= 0xBF58476D1CE4E5B9ULL;
        h ^= h >> 32;
        uint64_t v[4];
        ld32<V>(buf + (h % n_sectors) * 4, v);
        acc += v[0] ^ v[1] ^ v[2] ^ v[3];
    }
    if (acc == 0x1234567) *sink = acc;
}
template <int V>
void run(const char *name, const uint64_t *buf, uint64_t n_sectors, uint64_t n_loads, uint64_t *sink) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    k<V><<<148 * 8, 256>>>(buf, n_sectors, n_loads, sink);
    cudaEventRecord(a);
    for (int r = 0; r < 5; r++) k<V><<<148 * 8, 256>>>(buf, n_sectors, n_loads, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    ms /= 5;
    printf("%-62s %.3f ms  %.2f Gloads/s  %.0f GB/s of sectors\n", name, ms, n_loads / ms / 1e6, n_loads * 32 / ms / 1e6);
}
int main() {
    const uint64_t bytes = 1ull << 30, n_sectors = bytes / 32, n_loads = 1ull << 26;
    uint64_t *buf, *sink;
    cudaMalloc(&buf, bytes);
    cudaMalloc(&sink, 8);
    cudaMemset(buf, 1, bytes);
    size_t g = 0;
    cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
    printf("cudaLimitMaxL2FetchGranularity = %zu\n", g);
    run<0>("ld.global.nc.v4.u64", buf, n_sectors, n_loads, sink);
    run<1>("ld.global.nc.L2::evict_first.v4.u64", buf, n_sectors, n_loads, sink);
    run<2>("ld.global.nc.L2::64B.v4.u64", buf, n_sectors, n_loads, sink);
    run<3>("ld.global.nc.L2::128B.v4.u64", buf, n_sectors, n_loads, sink);
    run<4>("ld.global.nc.L1::no_allocate.v4.u64", buf, n_sectors, n_loads, sink);
    run<5>("2 x ld.global.nc.v2.u64", buf, n_sectors, n_loads, sink);
    run<6>("ld.global.cg.v4.u64", buf, n_sectors, n_loads, sink);
    run<7>("ld.global.nc.L1::no_allocate.L2::evict_first.L2::64B.v4.u64", buf, n_sectors, n_loads, sink);
    cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32);
    cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
    printf("cudaLimitMaxL2FetchGranularity = %zu\n", g);
    run<0>("ld.global.nc.v4.u64 (granularity 32)", buf, n_sectors, n_loads, sink);
    run<2>("ld.global.nc.L2::64B.v4.u64 (granularity 32)", buf, n_sectors, n_loads, sink);
    return 0;
}
