// np2_count.cu — `yak count` on the device (yak/count.c:28-165, yak/htab.c:51-78, yak/main.c:24-83): the producer of
// the k-mer tables this library probes.
//
// yak streams the reads through per-prefix hash tables (+ an optional Bloom filter pass).  What it leaves on disk is
// a pure function of the input: for every canonical k-mer hash its number of occurrences, saturated at 1023; with
// `-b N` (two passes + yak_ch_shrink(2, ..)) only the hashes seen at least twice.  Here that function is computed by
// sorting: hash every k-mer (one thread rolls over 32 consecutive end positions), radix-sort the hashes, run-length
// encode, merge with what earlier batches left (sort + reduce by key), clamp at 1023.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_reduce.cuh>
#include <cub/device/device_run_length_encode.cuh>
#include <cub/device/device_select.cuh>

#include <algorithm>
#include <string>

#include "../../include/np2gpu.h"
#include "np2_error.h"
#include "np2_kernels.cuh"

namespace np2 {

namespace {
inline uint32_t cdiv(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }
constexpr uint32_t kEndsPerThread = 32;

// count_seq_buf / count_seq_buf_long (yak/count.c:28-64): the rolling state machine, restarted by any non-ACGT byte.
// A thread owns the k-mer END positions [e0, e0 + 32) and warms its registers up on the k - 1 bases before e0.
__global__ void __launch_bounds__(256) k_count_hashes(const uint8_t *__restrict__ seq, uint64_t len, uint32_t k,
                                                      uint64_t *__restrict__ out, unsigned long long *__restrict__ n_out) {
    const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint64_t e0 = t * kEndsPerThread;
    const uint32_t lane = threadIdx.x & 31;
    uint64_t x0 = 0, x1 = 0, x2 = 0, x3 = 0;
    uint32_t l = 0;
    const bool small = k < 32;
    const uint64_t mask = small ? (1ULL << (2 * k)) - 1 : (1ULL << k) - 1;
    const uint32_t shift = small ? 2 * (k - 1) : k - 1;
    auto push = [&](uint32_t c) {
        if (c < 4) {
            if (small) {
                x0 = (x0 << 2 | c) & mask;
                x1 = x1 >> 2 | (uint64_t)(3 - c) << shift;
            } else {
                x0 = (x0 << 1 | (c & 1)) & mask;
                x1 = (x1 << 1 | (c >> 1)) & mask;
                x2 = x2 >> 1 | (uint64_t)(1 - (c & 1)) << shift;
                x3 = x3 >> 1 | (uint64_t)(1 - (c >> 1)) << shift;
            }
            l++;
        } else {
            l = 0;
            x0 = x1 = x2 = x3 = 0;
        }
    };
    if (e0 < len) {
        const uint64_t w0 = e0 >= k - 1 ? e0 - (k - 1) : 0;
        for (uint64_t p = w0; p < e0; p++) push(seq_code(seq[p]));
    }
    for (uint32_t i = 0; i < kEndsPerThread; i++) {  // every lane runs all 32 rounds (the ballot needs the whole warp)
        const uint64_t e = e0 + i;
        bool emit = false;
        uint64_t h = 0;
        if (e < len) {
            push(seq_code(seq[e]));
            if (l >= k) {
                emit = true;
                if (small) h = yak_hash64(x0 < x1 ? x0 : x1, mask);
                else h = x1 < x3 ? yak_hash64_64(x0) + yak_hash64_64(x1) : yak_hash64_64(x2) + yak_hash64_64(x3);
            }
        }
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, emit);
        if (bal) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(n_out, (unsigned long long)__popc(bal));
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (emit) out[base + __popc(bal & ((1u << lane) - 1))] = h;
        }
    }
}
__global__ void k_clamp_counts(uint32_t *__restrict__ c, uint64_t n) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n) c[i] = min(c[i], 1023u);
}
__global__ void k_count_flags(const uint32_t *__restrict__ c, uint64_t n, uint32_t min_count, uint8_t *__restrict__ flag,
                              const uint64_t *__restrict__ key, unsigned int *__restrict__ sub_size) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool keep = c[i] >= min_count;
    flag[i] = keep;
    if (keep) atomicAdd(&sub_size[key[i] & 1023], 1u);
}
__global__ void k_to_u16(const uint32_t *__restrict__ c, uint64_t n, uint16_t *__restrict__ o) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n) o[i] = (uint16_t)c[i];
}
// file key of yak's dump: (hash >> pre) << 10 | count, pre = 10 (yak/htab.c:59-66); sort key = sub-table
__global__ void k_file_keys(const uint64_t *__restrict__ key, const uint16_t *__restrict__ cnt, uint64_t n,
                            uint32_t *__restrict__ sub, uint64_t *__restrict__ fkey) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    sub[i] = (uint32_t)(key[i] & 1023);
    fkey[i] = (key[i] >> 10) << 10 | cnt[i];
}
template <class T>
T *dalloc(uint64_t n, cudaStream_t s) {
    T *p = nullptr;
    if (cudaMallocAsync((void **)&p, std::max<uint64_t>(n, 1) * sizeof(T), s) != cudaSuccess) {
        cudaGetLastError();
        throw Error(NP2_ERR_CUDA, "device allocation failed while counting k-mers");
    }
    return p;
}
#define NP2C(x)                                                                                      \
    do {                                                                                             \
        cudaError_t e_ = (x);                                                                        \
        if (e_ != cudaSuccess) throw Error(NP2_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_)); \
    } while (0)
}  // namespace

void count_free(KmerCounts &acc, cudaStream_t s) {
    if (acc.keys) cudaFreeAsync(acc.keys, s);
    if (acc.cnts) cudaFreeAsync(acc.cnts, s);
    acc = KmerCounts();
}

void count_add(KmerCounts &acc, const uint8_t *d_seq, uint64_t len, uint32_t k, uint64_t *n_kmers, cudaStream_t s) {
    if (len < k) return;
    if (len >= (1ull << 31)) throw Error(NP2_ERR_ARG, "k-mer counting batch must be below 2^31 bases");
    uint64_t *d_h = dalloc<uint64_t>(len, s), *d_h2 = dalloc<uint64_t>(len, s);
    unsigned long long *d_n = dalloc<unsigned long long>(1, s);
    NP2C(cudaMemsetAsync(d_n, 0, 8, s));
    NP2_K(k_count_hashes)<<<cdiv(cdiv(len, kEndsPerThread), 256), 256, 0, s>>>(d_seq, len, k, d_h, d_n);
    unsigned long long nv = 0;
    NP2C(cudaMemcpyAsync(&nv, d_n, 8, cudaMemcpyDeviceToHost, s));
    NP2C(cudaStreamSynchronize(s));
    if (n_kmers) *n_kmers += nv;
    void *d_tmp = nullptr;
    size_t tb = 0, cap = 0;
    auto need = [&](size_t b) {
        if (b > cap) {
            if (d_tmp) cudaFreeAsync(d_tmp, s);
            NP2C(cudaMallocAsync(&d_tmp, b, s));
            cap = b;
        }
    };
    uint64_t *d_uk = dalloc<uint64_t>(nv, s);
    uint32_t *d_uc = dalloc<uint32_t>(nv, s);
    uint32_t *d_runs = dalloc<uint32_t>(1, s);
    uint32_t runs = 0;
    if (nv) {
        const int end_bit = k < 32 ? (int)(2 * k) : 64;
        cub::DeviceRadixSort::SortKeys(nullptr, tb, d_h, d_h2, (int)nv, 0, end_bit, s);
        need(tb);
        cub::DeviceRadixSort::SortKeys(d_tmp, tb, d_h, d_h2, (int)nv, 0, end_bit, s);
        tb = 0;
        cub::DeviceRunLengthEncode::Encode(nullptr, tb, d_h2, d_uk, d_uc, d_runs, (int)nv, s);
        need(tb);
        cub::DeviceRunLengthEncode::Encode(d_tmp, tb, d_h2, d_uk, d_uc, d_runs, (int)nv, s);
        NP2C(cudaMemcpyAsync(&runs, d_runs, 4, cudaMemcpyDeviceToHost, s));
        NP2C(cudaStreamSynchronize(s));
    }
    cudaFreeAsync(d_h, s);
    cudaFreeAsync(d_h2, s);
    cudaFreeAsync(d_n, s);
    if (acc.n == 0) {
        count_free(acc, s);
        acc.keys = d_uk;
        acc.cnts = d_uc;
        acc.n = runs;
    } else if (runs) {  // merge: concatenate, sort by hash, add up
        const uint64_t m = acc.n + runs;
        if (m >= (1ull << 31)) throw Error(NP2_ERR_UNSUPPORTED, "more than 2^31 distinct k-mers in one counter");
        uint64_t *k1 = dalloc<uint64_t>(m, s), *k2 = dalloc<uint64_t>(m, s), *k3 = dalloc<uint64_t>(m, s);
        uint32_t *c1 = dalloc<uint32_t>(m, s), *c2 = dalloc<uint32_t>(m, s), *c3 = dalloc<uint32_t>(m, s);
        NP2C(cudaMemcpyAsync(k1, acc.keys, acc.n * 8, cudaMemcpyDeviceToDevice, s));
        NP2C(cudaMemcpyAsync(k1 + acc.n, d_uk, (uint64_t)runs * 8, cudaMemcpyDeviceToDevice, s));
        NP2C(cudaMemcpyAsync(c1, acc.cnts, acc.n * 4, cudaMemcpyDeviceToDevice, s));
        NP2C(cudaMemcpyAsync(c1 + acc.n, d_uc, (uint64_t)runs * 4, cudaMemcpyDeviceToDevice, s));
        const int end_bit = k < 32 ? (int)(2 * k) : 64;
        tb = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tb, k1, k2, c1, c2, (int)m, 0, end_bit, s);
        need(tb);
        cub::DeviceRadixSort::SortPairs(d_tmp, tb, k1, k2, c1, c2, (int)m, 0, end_bit, s);
        tb = 0;
        cub::DeviceReduce::ReduceByKey(nullptr, tb, k2, k3, c2, c3, d_runs, cub::Sum(), (int)m, s);
        need(tb);
        cub::DeviceReduce::ReduceByKey(d_tmp, tb, k2, k3, c2, c3, d_runs, cub::Sum(), (int)m, s);
        NP2C(cudaMemcpyAsync(&runs, d_runs, 4, cudaMemcpyDeviceToHost, s));
        NP2C(cudaStreamSynchronize(s));
        cudaFreeAsync(k1, s);
        cudaFreeAsync(k2, s);
        cudaFreeAsync(c1, s);
        cudaFreeAsync(c2, s);
        cudaFreeAsync(d_uk, s);
        cudaFreeAsync(d_uc, s);
        count_free(acc, s);
        acc.keys = k3;
        acc.cnts = c3;
        acc.n = runs;
    } else {
        cudaFreeAsync(d_uk, s);
        cudaFreeAsync(d_uc, s);
    }
    if (acc.n) NP2_K(k_clamp_counts)<<<cdiv(acc.n, 256), 256, 0, s>>>(acc.cnts, acc.n);  // counters stop at 1023 (htab.c:68-69)
    cudaFreeAsync(d_runs, s);
    if (d_tmp) cudaFreeAsync(d_tmp, s);
    NP2C(cudaStreamSynchronize(s));
}

void count_filter(const KmerCounts &acc, uint32_t min_count, uint64_t **d_keys, uint16_t **d_cnt, uint64_t *n,
                  uint32_t sub_size[1024], cudaStream_t s) {
    uint8_t *d_flag = dalloc<uint8_t>(acc.n, s);
    unsigned int *d_sub = dalloc<unsigned int>(1024, s);
    uint64_t *d_k = dalloc<uint64_t>(acc.n, s);
    uint32_t *d_c32 = dalloc<uint32_t>(acc.n, s);
    uint16_t *d_c16 = dalloc<uint16_t>(acc.n, s);
    uint32_t *d_num = dalloc<uint32_t>(1, s);
    NP2C(cudaMemsetAsync(d_sub, 0, 4096, s));
    uint32_t kept = 0;
    if (acc.n) {
        NP2_K(k_count_flags)<<<cdiv(acc.n, 256), 256, 0, s>>>(acc.cnts, acc.n, min_count, d_flag, acc.keys, d_sub);
        void *d_tmp = nullptr;
        size_t tb = 0, tb2 = 0;
        cub::DeviceSelect::Flagged(nullptr, tb, acc.keys, d_flag, d_k, d_num, (int)acc.n, s);
        cub::DeviceSelect::Flagged(nullptr, tb2, acc.cnts, d_flag, d_c32, d_num, (int)acc.n, s);
        tb = std::max(tb, tb2);
        NP2C(cudaMallocAsync(&d_tmp, tb, s));
        cub::DeviceSelect::Flagged(d_tmp, tb, acc.keys, d_flag, d_k, d_num, (int)acc.n, s);
        cub::DeviceSelect::Flagged(d_tmp, tb, acc.cnts, d_flag, d_c32, d_num, (int)acc.n, s);
        NP2C(cudaMemcpyAsync(&kept, d_num, 4, cudaMemcpyDeviceToHost, s));
        NP2C(cudaStreamSynchronize(s));
        if (kept) NP2_K(k_to_u16)<<<cdiv(kept, 256), 256, 0, s>>>(d_c32, kept, d_c16);
        cudaFreeAsync(d_tmp, s);
    }
    NP2C(cudaMemcpyAsync(sub_size, d_sub, 4096, cudaMemcpyDeviceToHost, s));
    NP2C(cudaStreamSynchronize(s));
    cudaFreeAsync(d_flag, s);
    cudaFreeAsync(d_sub, s);
    cudaFreeAsync(d_c32, s);
    cudaFreeAsync(d_num, s);
    *d_keys = d_k;
    *d_cnt = d_c16;
    *n = kept;
}

void count_file_keys(const uint64_t *d_keys, const uint16_t *d_cnt, uint64_t n, uint64_t *h_out, cudaStream_t s) {
    if (!n) return;
    uint32_t *d_sub = dalloc<uint32_t>(n, s), *d_sub2 = dalloc<uint32_t>(n, s);
    uint64_t *d_f = dalloc<uint64_t>(n, s), *d_f2 = dalloc<uint64_t>(n, s);
    NP2_K(k_file_keys)<<<cdiv(n, 256), 256, 0, s>>>(d_keys, d_cnt, n, d_sub, d_f);
    void *d_tmp = nullptr;
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, d_sub, d_sub2, d_f, d_f2, (int)n, 0, 10, s);
    NP2C(cudaMallocAsync(&d_tmp, tb, s));
    cub::DeviceRadixSort::SortPairs(d_tmp, tb, d_sub, d_sub2, d_f, d_f2, (int)n, 0, 10, s);
    NP2C(cudaMemcpyAsync(h_out, d_f2, n * 8, cudaMemcpyDeviceToHost, s));
    NP2C(cudaStreamSynchronize(s));
    cudaFreeAsync(d_sub, s);
    cudaFreeAsync(d_sub2, s);
    cudaFreeAsync(d_f, s);
    cudaFreeAsync(d_f2, s);
    cudaFreeAsync(d_tmp, s);
}

}  // namespace np2
