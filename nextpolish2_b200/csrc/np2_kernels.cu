// np2_kernels.cu — hand-written sm_100a kernels of the polish path.
//
//   K5  table_insert / table_probe / seq_kscore   yak k-mer table in HBM (kmer.rs:113-170, 255-314)
//   K0  ref_codes                                  SEQ_NUM codes of the contig (kmer.rs:11-22)
//   K1  trim_scan + pack_columns                        fill_with_cigar + trim(8) + AlignSeq::new (main.rs:386-513, 279-312)
//   K2  cover_diff / pileup_stripe                 update_msas + Msa::push/sort/coverage (main.rs:576-589, 193-241)
//   K3  dp_runs / emit_*                           get_cns_from_align_tags + backtrack (main.rs:1645-1687, 1572-1634)
//   (K4/K6 and the genotype kernels live in np2_geno.cu)
//
// All of this is integer / byte work bound by HBM traffic and latency; there is no tensor-core term.
#include <cuda/ptx>
#include <cub/block/block_reduce.cuh>
#include <cub/block/block_scan.cuh>

#include "np2_kernels.cuh"

namespace np2 {

unsigned long long &launch_counter() {
    static thread_local unsigned long long c = 0;  // a job runs on its caller's thread
    return c;
}

namespace {
constexpr int kThreads = 256;
inline uint32_t cdiv(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }
}  // namespace

/* =============================================================== K5: yak table */

__device__ __forceinline__ uint32_t bucket_of(uint64_t tag, uint32_t nb) {
    uint32_t m = (uint32_t)((tag * 0x9E3779B97F4A7C15ULL) >> 32);
    return (uint32_t)(((uint64_t)m * nb) >> 32);
}

__device__ __forceinline__ void insert_key(uint64_t *slots, uint32_t nb, uint32_t sub, uint64_t v, int *err) {
    const uint64_t tag = v >> 10;
    uint32_t b = bucket_of(tag, nb);
    for (uint32_t step = 0; step < nb; step++) {
        unsigned long long *bp = (unsigned long long *)(slots + ((uint64_t)sub * nb + b) * kBucketSlots);
#pragma unroll
        for (int j = 0; j < kBucketSlots; j++) {
            unsigned long long old = atomicCAS(bp + j, 0ULL, (unsigned long long)v);
            if (old == 0ULL || (old >> 10) == tag) return;
        }
        b = b + 1 == nb ? 0 : b + 1;
    }
    atomicExch(err, 1);
}

__global__ void k_table_insert(uint64_t *slots, uint32_t nb, const uint64_t *__restrict__ hashes,
                               const uint16_t *__restrict__ counts, uint64_t n, int *err) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t h = hashes[i];
        insert_key(slots, nb, (uint32_t)(h & 1023), (h >> 10) << 10 | (counts[i] & 1023), err);
    }
}

__global__ void k_table_insert_filekeys(uint64_t *slots, uint32_t nb, const uint64_t *__restrict__ keys,
                                        const uint32_t *__restrict__ sub_off, uint64_t n, int *err) {
    __shared__ uint32_t s_off[1025];
    for (int i = threadIdx.x; i < 1025; i += blockDim.x) s_off[i] = sub_off[i];
    __syncthreads();
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t lo = 0, hi = 1024;  // largest sub with s_off[sub] <= i
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (s_off[mid] <= i) lo = mid;
            else hi = mid;
        }
        insert_key(slots, nb, lo, keys[i], err);
    }
}

// One probe = one 32-byte sector: 8 B query in, 32 B bucket, 2 B out.  Each thread keeps kProbeIlp
// independent probes in flight so that DRAM latency is covered by memory-level parallelism.
// WIDE: the bucket comes in one 256-bit evict-first load (ld_bucket); !WIDE: two 128-bit __ldg (kept for A/B runs,
// NP2_PROBE_WIDE=0).
constexpr int kProbeIlp = 4;
template <bool WIDE>
__device__ __forceinline__ void load_bucket(const uint64_t *bp, uint64_t v[4]) {
    if (WIDE) {
        ld_bucket(bp, v);
    } else {
        const ulonglong2 lo = __ldg((const ulonglong2 *)bp), hi = __ldg((const ulonglong2 *)bp + 1);
        v[0] = lo.x, v[1] = lo.y, v[2] = hi.x, v[3] = hi.y;
    }
}
template <bool WIDE>
__device__ __forceinline__ uint16_t probe_finish(const uint64_t *__restrict__ slots, uint32_t nb, uint32_t sub,
                                                 uint32_t b, uint64_t tag, uint64_t v[4], uint32_t min_count) {
    for (uint32_t step = 0;; step++) {
        bool empty = false;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if ((v[j] >> 10) == tag && v[j] != 0) {
                uint32_t c = (uint32_t)(v[j] & 1023);
                return c >= min_count ? (uint16_t)c : (uint16_t)0;
            }
            empty |= (v[j] == 0);
        }
        if (empty || step + 1 >= nb) return 0;
        b = b + 1 == nb ? 0 : b + 1;
        load_bucket<WIDE>(slots + ((uint64_t)sub * nb + b) * kBucketSlots, v);
    }
}

template <bool WIDE>
__global__ void __launch_bounds__(kThreads) k_table_probe(const uint64_t *__restrict__ slots, uint32_t nb,
                                                          const uint64_t *__restrict__ hashes, uint64_t n,
                                                          uint32_t min_count, uint16_t *__restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t i0 = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    for (; i0 < n; i0 += stride * kProbeIlp) {
        uint64_t tag[kProbeIlp];
        uint32_t sub[kProbeIlp], b[kProbeIlp];
        uint64_t v[kProbeIlp][4];
#pragma unroll
        for (int u = 0; u < kProbeIlp; u++) {
            uint64_t i = i0 + u * stride;
            uint64_t h = i < n ? hashes[i] : 0;
            sub[u] = (uint32_t)(h & 1023);
            tag[u] = h >> 10;
            b[u] = bucket_of(tag[u], nb);
        }
#pragma unroll
        for (int u = 0; u < kProbeIlp; u++)
            load_bucket<WIDE>(slots + ((uint64_t)sub[u] * nb + b[u]) * kBucketSlots, v[u]);
#pragma unroll
        for (int u = 0; u < kProbeIlp; u++) {
            uint64_t i = i0 + u * stride;
            if (i < n) out[i] = probe_finish<WIDE>(slots, nb, sub[u], b[u], tag[u], v[u], min_count);
        }
    }
}

__device__ __forceinline__ uint16_t probe_one(const uint64_t *__restrict__ slots, uint32_t nb, uint64_t h,
                                              uint32_t min_count) {
    uint32_t sub = (uint32_t)(h & 1023);
    uint64_t tag = h >> 10;
    uint32_t b = bucket_of(tag, nb);
    uint64_t v[4];
    load_bucket<true>(slots + ((uint64_t)sub * nb + b) * kBucketSlots, v);
    return probe_finish<true>(slots, nb, sub, b, tag, v, min_count);
}

// kscore of a byte string: one warp per string, one lane per k-mer end position (iter2kmer kmer.rs:255-314,
// to_hash 102-110, min main.rs:761-769).  A window containing a non-ACGT code has no k-mer.
__global__ void __launch_bounds__(kThreads) k_seq_kscore(const uint64_t *__restrict__ slots, uint32_t nb, uint32_t k,
                                                         const uint8_t *__restrict__ seqs,
                                                         const uint64_t *__restrict__ off,
                                                         const uint32_t *__restrict__ sel, uint64_t n,
                                                         uint32_t min_count, uint16_t *__restrict__ out) {
    const uint32_t lane = threadIdx.x & 31;
    uint64_t w = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (; w < n; w += nw) {
        const uint64_t si = sel ? sel[w] : w;
        const uint8_t *s = seqs + off[si];
        const uint32_t len = (uint32_t)(off[si + 1] - off[si]);
        uint32_t mn = 0xFFFFFFFFu;
        for (uint32_t e = k - 1 + lane; e < len; e += 32) {  // k-mer occupying [e-k+1, e]
            bool ok = true;
            uint64_t hsh;
            if (k < 32) {
                const uint64_t mask = (1ULL << (2 * k)) - 1;
                uint64_t f = 0, r = 0;
                for (uint32_t x = e + 1 - k; x <= e; x++) {
                    uint32_t c = seq_code(s[x]);
                    ok &= c < 4;
                    f = (f << 2 | c) & mask;
                    r = (r >> 2) | (uint64_t)(3 ^ c) << (2 * (k - 1));
                }
                hsh = yak_hash64(f < r ? f : r, mask);
            } else {
                const uint64_t mask = (1ULL << k) - 1;
                uint64_t x0 = 0, x1 = 0, x2 = 0, x3 = 0;
                for (uint32_t x = e + 1 - k; x <= e; x++) {
                    uint64_t c = seq_code(s[x]);
                    ok &= c < 4;
                    x0 = (x0 << 1 | (c & 1)) & mask;
                    x1 = (x1 << 1 | (c >> 1)) & mask;
                    x2 = x2 >> 1 | (1 - (c & 1)) << (k - 1);
                    x3 = x3 >> 1 | (1 - (c >> 1)) << (k - 1);
                }
                hsh = x1 < x3 ? yak_hash64_64(x0) + yak_hash64_64(x1) : yak_hash64_64(x2) + yak_hash64_64(x3);
            }
            if (ok) mn = min(mn, (uint32_t)probe_one(slots, nb, hsh, min_count));
        }
        mn = __reduce_min_sync(0xFFFFFFFFu, mn);
        if (lane == 0) out[w] = mn == 0xFFFFFFFFu ? 0 : (uint16_t)mn;
    }
}

void table_insert(const TableDev &t, const uint64_t *d_hashes, const uint16_t *d_counts, uint64_t n, int *d_err,
                  cudaStream_t s) {
    if (!n) return;
    NP2_K(k_table_insert)<<<min(cdiv(n, kThreads), 148u * 32u), kThreads, 0, s>>>(t.slots, t.nb, d_hashes, d_counts, n, d_err);
}
void table_insert_filekeys(const TableDev &t, const uint64_t *d_keys, const uint32_t *d_sub_off, uint64_t n,
                           int *d_err, cudaStream_t s) {
    if (!n) return;
    NP2_K(k_table_insert_filekeys)<<<min(cdiv(n, kThreads), 148u * 32u), kThreads, 0, s>>>(t.slots, t.nb, d_keys, d_sub_off, n,
                                                                                   d_err);
}
void table_probe(const TableDev &t, const uint64_t *d_hashes, uint64_t n, uint32_t min_count, uint16_t *d_out,
                 cudaStream_t s) {
    if (!n) return;
    // persistent-style grid: a multiple of the 148 SMs x 8 resident CTAs of 256 threads
    uint32_t grid = min(cdiv(n, (uint64_t)kThreads * kProbeIlp), 148u * 8u);
    static const bool wide = [] {
        const char *e = getenv("NP2_PROBE_WIDE");
        return !e || atoi(e) != 0;
    }();
    if (wide) NP2_K(k_table_probe<true>)<<<grid, kThreads, 0, s>>>(t.slots, t.nb, d_hashes, n, min_count, d_out);
    else NP2_K(k_table_probe<false>)<<<grid, kThreads, 0, s>>>(t.slots, t.nb, d_hashes, n, min_count, d_out);
}
void seq_kscore(const TableDev &t, const uint8_t *d_seqs, const uint64_t *d_off, const uint32_t *d_sel, uint64_t n,
                uint32_t min_count, uint16_t *d_out, cudaStream_t s) {
    if (!n) return;
    uint32_t grid = min(cdiv(n * 32, kThreads), 148u * 8u);
    NP2_K(k_seq_kscore)<<<grid, kThreads, 0, s>>>(t.slots, t.nb, t.k, d_seqs, d_off, d_sel, n, min_count, d_out);
}

/* --------------------------------------------------------------- measurement: random 32-B sector gather
 * The denominator for K5's roofline: independent uniformly random 32-byte sector reads over a buffer far larger
 * than L2, same ILP and grid shape as k_table_probe but no hashing, no compare, no dependent second probe. */
// BYTES = 32: one 256-bit load per random sector; 64 / 128: the 2 / 4 sectors of an aligned 64- / 128-byte block, for the
// comparison that shows what granularity DRAM really serves (profiles/)
template <int BYTES>
__global__ void __launch_bounds__(kThreads) k_gather32(const uint64_t *__restrict__ buf, uint64_t n_sectors,
                                                       uint64_t n_loads, uint64_t seed, uint64_t *__restrict__ sink) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    constexpr int kS = BYTES / 32;
    uint64_t acc = 0;
    for (uint64_t i0 = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i0 < n_loads; i0 += stride * kProbeIlp) {
        uint64_t v[kProbeIlp][kS][4];
#pragma unroll
        for (int u = 0; u < kProbeIlp; u++) {
            uint64_t x = (i0 + u * stride + seed) * 0x9E3779B97F4A7C15ULL;
            x ^= x >> 29;
            x *= 0xBF58476D1CE4E5B9ULL;
            x ^= x >> 32;
            const uint64_t sct = (uint64_t)(((unsigned __int128)x * (n_sectors / kS)) >> 64) * kS;
#pragma unroll
            for (int q = 0; q < kS; q++) ld_bucket(buf + (sct + q) * 4, v[u][q]);
        }
#pragma unroll
        for (int u = 0; u < kProbeIlp; u++)
#pragma unroll
            for (int q = 0; q < kS; q++) acc ^= v[u][q][0] ^ v[u][q][1] ^ v[u][q][2] ^ v[u][q][3];
    }
    if (acc == 0x123456789ABCDEFULL) *sink = acc;  // keeps the loads alive
}
void gather32(const uint64_t *d_buf, uint64_t n_sectors, uint64_t n_loads, uint64_t seed, uint64_t *d_sink,
              cudaStream_t s, int bytes) {
    uint32_t grid = min(cdiv(n_loads, (uint64_t)kThreads * kProbeIlp), 148u * 8u);
    if (bytes == 128) NP2_K(k_gather32<128>)<<<grid, kThreads, 0, s>>>(d_buf, n_sectors, n_loads, seed, d_sink);
    else if (bytes == 64) NP2_K(k_gather32<64>)<<<grid, kThreads, 0, s>>>(d_buf, n_sectors, n_loads, seed, d_sink);
    else NP2_K(k_gather32<32>)<<<grid, kThreads, 0, s>>>(d_buf, n_sectors, n_loads, seed, d_sink);
}

/* =============================================================== K0: reference codes */

// also flags the bytes the reference cannot index (>= 128 panics in SEQ_NUM[..]; '-' would read as a gap column)
__global__ void k_ref_codes(const uint8_t *__restrict__ ref, uint32_t L, uint8_t *__restrict__ code, int *bad) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L) return;
    const uint32_t c = ref[i];
    code[i] = (uint8_t)seq_code(c);
    if (c >= 128 || c == '-') atomicExch(bad, 1);
}
// 8 codes per u32, position 8w in the most significant nibble: a read block's 32 plain columns can then be compared
// with the reference in four word operations (scan_block32 fast path).  The array is padded with 0xF nibbles.
__global__ void k_ref_pack(const uint8_t *__restrict__ code, uint32_t L, uint32_t n_words, uint32_t *__restrict__ pk) {
    uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    uint32_t v = 0;
#pragma unroll
    for (uint32_t x = 0; x < 8; x++) {
        const uint32_t p = w * 8 + x;
        v = v << 4 | (p < L ? code[p] : 15u);
    }
    pk[w] = v;
}
void ref_codes(const uint8_t *d_ref, uint32_t L, uint8_t *d_code, uint32_t *d_refpk, int *d_bad, cudaStream_t s) {
    NP2_K(k_ref_codes)<<<cdiv(L, kThreads), kThreads, 0, s>>>(d_ref, L, d_code, d_bad);
    const uint32_t nw = L / 8 + 8;
    NP2_K(k_ref_pack)<<<cdiv(nw, kThreads), kThreads, 0, s>>>(d_code, L, nw, d_refpk);
}

/* =============================================================== K0: SEQ gather from the caller's pinned records */
// The BAM records of a contig are ~2/3 QUAL bytes, read names and tags the path never looks at.  When the caller's
// record buffer is page-locked, the device pulls just the 4-bit SEQ fields over PCIe itself (mapped host memory, UVA):
// one CTA per read, 16-byte vectors, source and destination share the same misalignment so every access is aligned.
// A vector that straddles the SEQ ends only touches bytes of the same 16-byte line, i.e. of a mapped page.
__global__ void __launch_bounds__(256) k_gather_seq(const uint8_t *__restrict__ src, const uint64_t *__restrict__ src_off,
                                                    const uint64_t *__restrict__ dst_off,
                                                    const uint32_t *__restrict__ nbytes, uint8_t *__restrict__ dst,
                                                    uint32_t n_reads) {
    for (uint32_t r = blockIdx.x; r < n_reads; r += gridDim.x) {
        const uint8_t *a = src + src_off[r];
        const uint32_t mis = (uint32_t)((uintptr_t)a & 15);
        const uint4 *sp = reinterpret_cast<const uint4 *>(a - mis);
        uint4 *dp = reinterpret_cast<uint4 *>(dst + dst_off[r] - mis);
        const uint32_t nv = (mis + nbytes[r] + 15) >> 4;
        for (uint32_t i = threadIdx.x; i < nv; i += 4 * 256) {
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (i + u * 256 < nv) v[u] = __ldcs(sp + i + u * 256);
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (i + u * 256 < nv) dp[i + u * 256] = v[u];
        }
    }
}
// ---- the same gather with the TMA engines (cp.async.bulk, 1-D): one thread per CTA drives a two-stage pipeline
// host memory -> shared memory -> device memory; no register, LSU or warp-slot cost per byte, so the SMs stay free
// for the kernels of the other contigs in flight.  Source, destination and size are multiples of 16 bytes by
// construction (the destination slot keeps the source's misalignment).
constexpr uint32_t kTmaStage = 16u << 10;  // 2 stages x 16 KB per CTA: a small shared-memory footprint keeps the L1 of
                                           // the kernels running next to it
__global__ void __launch_bounds__(32) k_gather_seq_tma(const uint8_t *__restrict__ src, const uint64_t *__restrict__ src_off,
                                                       const uint64_t *__restrict__ dst_off,
                                                       const uint32_t *__restrict__ nbytes, uint8_t *__restrict__ dst,
                                                       uint32_t n_reads) {
    extern __shared__ __align__(128) uint8_t smem[];  // 2 stages
    __shared__ __align__(8) uint64_t bar[2];
    if (threadIdx.x != 0) return;
    namespace ptx = cuda::ptx;
    ptx::mbarrier_init(&bar[0], 1);
    ptx::mbarrier_init(&bar[1], 1);
    ptx::fence_proxy_async(ptx::space_shared);  // barrier initialisation visible to the async proxy
    uint32_t phase[2] = {0, 0};
    uint32_t issued = 0;       // chunks whose load has been issued
    uint8_t *pend_dst = nullptr;  // the chunk loaded into stage (issued - 1) & 1, waiting to be stored
    uint32_t pend_bytes = 0;
    auto flush = [&]() {  // wait for the pending load, store it
        if (!pend_bytes) return;
        const uint32_t st = (issued - 1) & 1;
        while (!ptx::mbarrier_try_wait_parity(&bar[st], phase[st])) {
        }
        phase[st] ^= 1;
        ptx::cp_async_bulk(ptx::space_global, ptx::space_shared, pend_dst, smem + st * kTmaStage, pend_bytes);
        ptx::cp_async_bulk_commit_group();
        pend_bytes = 0;
    };
    for (uint32_t r = blockIdx.x; r < n_reads; r += gridDim.x) {
        const uint8_t *a = src + src_off[r];
        const uint32_t mis = (uint32_t)((uintptr_t)a & 15);
        const uint8_t *sp = a - mis;
        uint8_t *dp = dst + dst_off[r] - mis;
        const uint32_t total = ((mis + nbytes[r] + 15) >> 4) << 4;
        for (uint32_t o = 0; o < total; o += kTmaStage) {
            const uint32_t bytes = min(kTmaStage, total - o);
            const uint32_t st = issued & 1;
            // the store that last read this stage was committed in the previous round: it must have finished READING
            // shared memory (not writing global memory) before the stage is loaded again
            ptx::cp_async_bulk_wait_group_read(ptx::n32_t<0>());
            ptx::mbarrier_arrive_expect_tx(ptx::sem_release, ptx::scope_cta, ptx::space_shared, &bar[st], bytes);
            ptx::cp_async_bulk(ptx::space_cluster, ptx::space_global, smem + st * kTmaStage, sp + o, bytes, &bar[st]);
            uint8_t *this_dst = dp + o;
            issued++;
            // while this load is in flight, retire the previous one
            if (pend_bytes) {
                const uint32_t pst = (issued - 2) & 1;
                while (!ptx::mbarrier_try_wait_parity(&bar[pst], phase[pst])) {
                }
                phase[pst] ^= 1;
                ptx::cp_async_bulk(ptx::space_global, ptx::space_shared, pend_dst, smem + pst * kTmaStage, pend_bytes);
                ptx::cp_async_bulk_commit_group();
            }
            pend_dst = this_dst;
            pend_bytes = bytes;
        }
    }
    flush();
    ptx::cp_async_bulk_wait_group(ptx::n32_t<0>());
}
// A small persistent grid (two CTAs per SM): the loads wait on PCIe, not on the SMs, and ~2.4 MB in flight saturate
// the link; a CTA per read would fill every SM with waiting threads and lock out the kernels of the other contig
// that is in flight on this GPU.
void gather_seq(const uint8_t *src_mapped, const uint64_t *d_src_off, const uint64_t *d_dst_off, const uint32_t *d_nbytes,
                uint8_t *d_dst, uint32_t n_reads, cudaStream_t s, bool src_on_device) {
    static const uint32_t ctas = getenv("NP2_K0_CTAS") ? (uint32_t)atoi(getenv("NP2_K0_CTAS")) : 2 * 148;
    static const bool use_tma = !(getenv("NP2_K0_TMA") && atoi(getenv("NP2_K0_TMA")) == 0);
    if (n_reads && use_tma) {
        static bool attr_set = false;
        if (!attr_set) {
            cudaFuncSetAttribute(k_gather_seq_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * kTmaStage));
            attr_set = true;
        }
        // 16 single-thread CTAs keep the link full (two 16 KB stages each); one per SM (148) measured the same alone and
        // 8 % slower with contigs in flight: a resident CTA pins its SM's shared-memory configuration for milliseconds
        static const uint32_t tma_ctas = getenv("NP2_K0_CTAS") ? ctas : 16;
        // records already on this device (np2_job_create_bgzf): the gather runs against HBM, not the link — a CTA pair
        // per SM keeps enough 16 KB bulk copies in flight (16 CTAs: 0.80 ms for 165 MB, profiles/r02bc_bgzf_job_launches.txt)
        const uint32_t want = src_on_device ? 2 * 148 : tma_ctas;
        NP2_K(k_gather_seq_tma)<<<std::min<uint32_t>(n_reads, std::max(1u, want)), 32, 2 * kTmaStage, s>>>(
            src_mapped, d_src_off, d_dst_off, d_nbytes, d_dst, n_reads);
        return;
    }
    if (n_reads)
        NP2_K(k_gather_seq)<<<std::min<uint32_t>(n_reads, std::max(1u, ctas)), 256, 0, s>>>(src_mapped, d_src_off, d_dst_off,
                                                                                       d_nbytes, d_dst, n_reads);
}

/* =============================================================== K1: expand + trim + pack */

struct OpCur {
    uint32_t i, i_end;        // current op, one past the read's last op
    uint32_t c_beg, c_end;    // columns covered by the op
    uint32_t q, t, op;
};
__device__ __forceinline__ void op_load(const ReadsDev &R, OpCur &c) {
    const uint4 o = __ldg(R.ops + c.i);
    c.op = o.w & 15;
    c.c_beg = o.x;
    c.c_end = o.x + (o.w >> 4);
    c.q = o.y;
    c.t = o.z;
}
// position the cursor on the op containing column col (col < total columns)
__device__ __forceinline__ void op_seek(const ReadsDev &R, uint32_t r, uint32_t col, OpCur &c) {
    uint32_t lo = R.op_off[r], hi = R.op_off[r + 1];
    c.i_end = hi;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (R.ops[mid].x <= col) lo = mid;
        else hi = mid;
    }
    c.i = lo;
    op_load(R, c);
}
// the op holding the first column of block g (written by k_trim_scan); reads with >= 65535 ops search instead
__device__ __forceinline__ void op_at_block(const ReadsDev &R, uint32_t r, uint32_t g, uint32_t col, OpCur &c) {
    const uint32_t v = R.blk_op[g];
    if (v == 0xFFFFu) {
        op_seek(R, r, col, c);
        return;
    }
    c.i = R.op_off[r] + v;
    c.i_end = R.op_off[r + 1];
    op_load(R, c);
}
// column -> (match?, nibble).  M/=/X compare raw bytes like trim() (main.rs:454); the nibble is
// SEQ_NUM[q] | 8 for an insertion column (main.rs:292-294).
__device__ __forceinline__ void col_eval(const ReadsDev &R, const uint8_t *__restrict__ ref, uint32_t pos,
                                         const uint8_t *__restrict__ seq4, const OpCur &c, uint32_t col, bool &match,
                                         uint32_t &nib) {
    uint32_t off = col - c.c_beg;
    if (c.op == 2) {  // D
        match = false;
        nib = 4;
        return;
    }
    uint32_t qi = c.q + off;
    uint32_t byte = seq4[qi >> 1];
    uint32_t b4 = (qi & 1) ? (byte & 15) : (byte >> 4);
    if (c.op == 1) {  // I
        match = false;
        nib = bam4_code(b4) | 8;
    } else {
        match = ref[pos + c.t + off] == bam4_char(b4);
        nib = bam4_code(b4);
    }
}
__device__ __forceinline__ uint32_t block_mask16(const ReadsDev &R, const uint8_t *ref, uint32_t r, uint32_t pos,
                                                 const uint8_t *seq4, uint32_t blk, uint32_t C) {
    uint32_t c0 = blk * 16;
    if (c0 >= C) return 0;
    OpCur cur;
    op_seek(R, r, c0, cur);
    uint32_t m = 0;
    uint32_t c1 = min(c0 + 16, C);
    for (uint32_t c = c0; c < c1; c++) {
        while (c >= cur.c_end) {
            cur.i++;
            op_load(R, cur);
        }
        bool match;
        uint32_t nib;
        col_eval(R, ref, pos, seq4, cur, c, match, nib);
        m |= (uint32_t)match << (c - c0);
    }
    return m;
}
__device__ __forceinline__ uint32_t run8_starts(uint32_t x) {  // bit j set iff x[j..j+7] are all ones
    x &= x >> 1;
    x &= x >> 2;
    x &= x >> 4;
    return x;
}
// (t_pos, delta) of a column, from its op (get_align_tag main.rs:314-338 read forwards gives the same)
__device__ __forceinline__ void col_tpos(const ReadsDev &R, uint32_t r, uint32_t pos, const OpCur &c, uint32_t col,
                                         uint32_t &tpos, uint32_t &delta) {
    uint32_t off = col - c.c_beg;
    if (c.op == 1) {
        tpos = pos + c.t - 1;
        delta = off + 1;
        uint32_t i = c.i;
        while (i > R.op_off[r] && (R.ops[i - 1].w & 15) == 1) {
            delta += R.ops[i - 1].w >> 4;
            i--;
        }
    } else {
        tpos = pos + c.t + off;
        delta = 0;
    }
}

// One warp per read.
__global__ void __launch_bounds__(128) k_trim_scan(ReadsDev R, const uint8_t *__restrict__ ref, uint32_t L) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= R.n_reads) return;
    const uint32_t pos = R.pos[r];
    const uint32_t C = R.ncols[r];
    const uint8_t *seq4 = R.blob + R.seq_off[r];
    const uint32_t nb16 = (C + 15) >> 4;
    const uint32_t ck0 = R.ck_off[r], nck = R.ck_off[r + 1] - ck0;
    for (uint32_t b = lane; b < nck; b += 32) R.ck_read[ck0 + b] = r;

    // ---- trim start: first run of 8 matching columns (main.rs:453-476)
    uint32_t shift = 0xFFFFFFFFu;
    for (uint32_t base = 0; base < nb16; base += 31) {
        uint32_t m = block_mask16(R, ref, r, pos, seq4, base + lane, C);
        uint32_t nxt = __shfl_down_sync(0xFFFFFFFFu, m, 1);
        uint32_t st = lane < 31 ? (run8_starts(m | nxt << 16) & 0xFFFFu) : 0;
        uint32_t bal = __ballot_sync(0xFFFFFFFFu, st != 0);
        if (bal) {
            uint32_t src = __ffs(bal) - 1;
            uint32_t st0 = __shfl_sync(0xFFFFFFFFu, st, src);
            shift = (base + src) * 16 + (__ffs(st0) - 1);
            break;
        }
    }
    if (shift == 0xFFFFFFFFu) {  // no anchor: shift = len, aln_len = 0 (main.rs:510-512)
        if (lane == 0) {
            R.n[r] = 0;
            R.t_s[r] = pos;
            R.t_e[r] = pos;
        }
        return;
    }
    // ---- trim end: last run of 8 (main.rs:479-509)
    uint32_t new_len = 0;
    for (int64_t base = (int64_t)nb16 - 31;; base -= 31) {
        int64_t blk = base + lane;
        uint32_t m = (blk >= 0 && blk < (int64_t)nb16) ? block_mask16(R, ref, r, pos, seq4, (uint32_t)blk, C) : 0;
        uint32_t nxt = __shfl_down_sync(0xFFFFFFFFu, m, 1);
        uint32_t st = lane < 31 ? (run8_starts(m | nxt << 16) & 0xFFFFu) : 0;
        uint32_t bal = __ballot_sync(0xFFFFFFFFu, st != 0);
        if (bal) {
            uint32_t src = 31 - __clz(bal);
            uint32_t st0 = __shfl_sync(0xFFFFFFFFu, st, src);
            new_len = (uint32_t)(base + src) * 16 + (31 - __clz(st0)) + 8;
            break;
        }
        if (base + 30 < 0) break;
    }
    const uint32_t n = new_len - shift;
    if (lane == 0) {
        OpCur cur;
        uint32_t tp, dl;
        op_seek(R, r, shift, cur);
        col_tpos(R, r, pos, cur, shift, tp, dl);
        R.t_s[r] = tp;
        op_seek(R, r, new_len - 1, cur);
        col_tpos(R, r, pos, cur, new_len - 1, tp, dl);
        R.t_e[r] = tp;
        R.n[r] = n;
    }
    // the packing itself is a separate, perfectly parallel kernel (k_pack_columns); when the terminator falls on a
    // 32-column boundary no pack thread owns its word, so it is written here
    if (lane == 0) {
        R.shift[r] = shift;
        if ((n & 31) == 0) *(uint64_t *)(R.nib + R.nib_off[r] + (n >> 1)) = 0xFFFFFFFFFFFFFFFFULL;
    }
    // which op holds the first column of every 32-column block of the trimmed read: one lane per op writes the
    // (few, consecutive) blocks that start inside it, so that the pack threads need no search of their own
    const uint32_t op0 = R.op_off[r], op1 = R.op_off[r + 1], nblk = (n + 31) >> 5;
    for (uint32_t i = op0 + lane; i < op1; i += 32) {
        const uint4 o = __ldg(R.ops + i);
        const uint32_t c_end = o.x + (o.w >> 4);
        if (c_end <= shift) continue;
        const uint32_t b_lo = o.x > shift ? (o.x - shift + 31) >> 5 : 0;
        const uint32_t b_hi = min(nblk, ((c_end - 1 - shift) >> 5) + 1);  // exclusive
        const uint16_t v = (uint16_t)min(i - op0, 0xFFFFu);
        for (uint32_t b = b_lo; b < b_hi; b++) R.blk_op[ck0 + b] = v;
    }
}

// ---- pack columns [shift, shift + n) into nibbles + terminator (main.rs:287-310) and write the checkpoints.
// One thread per 32-column output block (16 bytes).  Blocks that lie inside one M/=/X op (~95 %) are 32 consecutive
// SEQ nibbles mapped through the 16-entry code table; the others (op boundaries, indels, the terminator) are queued
// in shared memory and packed column by column by densely filled warps.
constexpr int kPackThreads = 256;
constexpr int kPackBatchDefault = 2;  // blocks per thread of k_pack_columns_batched (1 = k_pack_columns)
// 16 BAM nibbles (first column in the top nibble) -> 16 SEQ_NUM codes in the byte order of memory (column 0 in the high
// nibble of byte 0).  Four nibbles at a time through PRMT: the 16-bit group IS the selector; two 8-entry byte tables
// (codes of "=ACMGRSV" and of "TWYHKDBN") are looked up with the low three bits, and a third PRMT over 0x80 bytes turns
// bit 3 of every nibble into a byte mask (sign replication) that picks between them.
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t c) {  // PTX generic form: selector bit 3 of
    uint32_t d;                                                                  // a nibble replicates the byte's msb
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t map_codes4(uint32_t sel) {
    const uint32_t s7 = sel & 0x7777u;
    const uint32_t lo = prmt(0x06010004u, 0x04040402u, s7);  // = A C M | G R S V -> 4 0 1 6 | 2 4 4 4
    const uint32_t hi = prmt(0x04040403u, 0x05040404u, s7);  // T W Y H | K D B N -> 3 4 4 4 | 4 4 4 5
    const uint32_t m = prmt(0x80808080u, 0x80808080u, sel);  // 0xFF where the nibble is >= 8, else 0x80
    const uint32_t r = (lo & ~m) | (hi & m);                        // one code per byte, last column in byte 0
    const uint32_t p = (r | r >> 4) & 0x00FF00FFu;                  // pairs: byte 0 = cols 2,3; byte 2 = cols 0,1
    return __byte_perm(p, 0, 0x4402);                               // 16 bits, cols 0,1 in the low (first) byte
}
__device__ __forceinline__ uint64_t map_codes16(uint64_t v) {
    return (uint64_t)map_codes4((uint32_t)(v >> 48)) | (uint64_t)map_codes4((uint32_t)(v >> 32)) << 16 |
           (uint64_t)map_codes4((uint32_t)(v >> 16)) << 32 | (uint64_t)map_codes4((uint32_t)v) << 48;
}
// 32 SEQ nibbles starting at query index qi, numeric order (first nibble at the top of hi).  Reads 20 bytes from an
// arbitrary address as five aligned words (every read slot has 32 bytes of slack behind it).
__device__ __forceinline__ void load_seq32(const uint8_t *__restrict__ seq4, uint32_t qi, uint64_t &hi, uint64_t &lo) {
    const uint8_t *sp = seq4 + (qi >> 1);
    const uint32_t mis = (uint32_t)((uintptr_t)sp & 3);
    const uint32_t *wp = (const uint32_t *)(sp - mis);
    const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3], w4 = wp[4];
    const uint32_t sh = mis * 8;
    // big-endian numeric words: first SEQ byte in the most significant position
    const uint32_t b0 = __byte_perm(__funnelshift_r(w0, w1, sh), 0, 0x0123);
    const uint32_t b1 = __byte_perm(__funnelshift_r(w1, w2, sh), 0, 0x0123);
    const uint32_t b2 = __byte_perm(__funnelshift_r(w2, w3, sh), 0, 0x0123);
    const uint32_t b3 = __byte_perm(__funnelshift_r(w3, w4, sh), 0, 0x0123);
    hi = (uint64_t)b0 << 32 | b1;
    lo = (uint64_t)b2 << 32 | b3;
    if (qi & 1) {  // odd query index: the stream starts at the low nibble of the first byte
        const uint32_t b16 = (__funnelshift_r(w4, 0, sh)) & 255u;  // 17th byte
        hi = hi << 4 | lo >> 60;
        lo = lo << 4 | (b16 >> 4);
    }
}
// 128-bit helpers on (hi, lo) = columns 0-15 / 16-31, column 0 in the top nibble of hi
__device__ __forceinline__ void shr_nib(uint64_t &hi, uint64_t &lo, uint32_t s) {  // towards later columns, s < 32
    if (s >= 16) {
        lo = hi >> (4 * (s - 16));
        hi = 0;
    } else if (s) {
        lo = lo >> (4 * s) | hi << (64 - 4 * s);
        hi >>= 4 * s;
    }
}
__device__ __forceinline__ void first_nib(uint32_t len, uint64_t &hi, uint64_t &lo) {  // 0xF in the first len columns
    hi = len >= 16 ? ~0ULL : (len ? ~0ULL << (64 - 4 * len) : 0);
    lo = len >= 32 ? ~0ULL : (len > 16 ? ~0ULL << (64 - 4 * (len - 16)) : 0);
}
// A block that is not one plain run of M/=/X columns (an op boundary, an indel, the end of the read): assembled
// segment by segment with the same word operations as the fast path instead of column by column.
__device__ __forceinline__ void pack_block_slow(const ReadsDev &R, uint32_t g) {
    const uint32_t r = R.ck_read[g];
    const uint32_t n = R.n[r], shift = R.shift[r];
    const uint32_t o0 = (g - R.ck_off[r]) * 32;
    const uint8_t *seq4 = R.blob + R.seq_off[r];
    const uint32_t ncol = min(32u, n > o0 ? n - o0 : 0u);  // columns of the read inside this block
    uint64_t src_hi = 0, src_lo = 0;                        // BAM nibbles of the M/=/X/I columns
    uint64_t ins_hi = 0, ins_lo = 0, del_hi = 0, del_lo = 0;  // 0xF where the column is an insertion / a deletion
    if (ncol) {
        OpCur cur;
        op_at_block(R, r, g, shift + o0, cur);
        uint32_t c = 0;  // column inside the block
        for (;;) {
            const uint32_t col = shift + o0 + c;
            const uint32_t len = min(cur.c_end - col, ncol - c);
            uint64_t mh, ml;
            first_nib(len, mh, ml);
            shr_nib(mh, ml, c);
            if (cur.op == 2) {
                del_hi |= mh;
                del_lo |= ml;
            } else {
                uint64_t h, l;
                load_seq32(seq4, cur.q + (col - cur.c_beg), h, l);
                shr_nib(h, l, c);
                src_hi |= h & mh;
                src_lo |= l & ml;
                if (cur.op == 1) {
                    ins_hi |= mh;
                    ins_lo |= ml;
                }
            }
            c += len;
            if (c >= ncol) break;
            cur.i++;
            op_load(R, cur);
        }
    }
    uint64_t th, tl;  // terminator / padding: 0xF in every column at or behind n
    first_nib(ncol, th, tl);
    th = ~th;
    tl = ~tl;
    const uint64_t k4 = 0x4444444444444444ULL, k8 = 0x8888888888888888ULL;
    // numeric nibble order -> memory order: reverse the bytes of each half
    auto to_mem = [](uint64_t x) {
        return (uint64_t)__byte_perm((uint32_t)(x >> 32), 0, 0x0123) | (uint64_t)__byte_perm((uint32_t)x, 0, 0x0123) << 32;
    };
    const uint64_t w0 = (map_codes16(src_hi) & ~to_mem(del_hi | th)) | to_mem((ins_hi & k8) | (del_hi & k4) | th);
    const uint64_t w1 = (map_codes16(src_lo) & ~to_mem(del_lo | tl)) | to_mem((ins_lo & k8) | (del_lo & k4) | tl);
    uint64_t *out = (uint64_t *)(R.nib + R.nib_off[r] + (o0 >> 1));
    out[0] = w0;
    out[1] = w1;
}
__global__ void __launch_bounds__(kPackThreads, 8) k_pack_columns(ReadsDev R, const uint8_t *__restrict__ ref,
                                                               uint32_t n_blocks) {
    __shared__ uint32_t q[kPackThreads], qn;
    if (threadIdx.x == 0) qn = 0;
    __syncthreads();
    const uint32_t g = blockIdx.x * kPackThreads + threadIdx.x;
    if (g < n_blocks) {
        const uint32_t r = R.ck_read[g];
        const uint32_t n = R.n[r];
        const uint32_t o0 = (g - R.ck_off[r]) * 32;
        if (n && o0 <= n) {
            bool done = false;
            if (o0 < n) {
                const uint32_t shift = R.shift[r], pos = R.pos[r];
                OpCur cur;
                op_at_block(R, r, g, shift + o0, cur);
                uint32_t tp, dl;
                col_tpos(R, r, pos, cur, shift + o0, tp, dl);
                R.ck_tpos[g] = tp;
                R.ck_delta[g] = (uint16_t)dl;
                if (o0 + 32 <= n && cur.op != 1 && cur.op != 2 && shift + o0 + 32 <= cur.c_end) {
                    uint64_t hi, lo;
                    load_seq32(R.blob + R.seq_off[r], cur.q + (shift + o0 - cur.c_beg), hi, lo);
                    uint64_t *out = (uint64_t *)(R.nib + R.nib_off[r] + (o0 >> 1));
                    out[0] = map_codes16(hi);
                    out[1] = map_codes16(lo);
                    done = true;
                }
            }
            if (!done) q[atomicAdd(&qn, 1u)] = g;
        }
    }
    __syncthreads();
    const uint32_t nq = qn;
    for (uint32_t i = threadIdx.x; i < nq; i += kPackThreads) pack_block_slow(R, q[i]);
}
// Two (B) blocks per thread with the loads of each level of the chain block -> read -> op -> SEQ words issued for all
// of them before the first use: one block per thread leaves too few bytes in flight per SM (56% of the samples of
// k_pack_columns were long-scoreboard stalls on that chain, profiles/r01end).  Plain blocks (32 M/=/X columns of one
// op) are finished here; everything else goes through the queue to the general per-block code.
__device__ __forceinline__ void pack_block_queued(const ReadsDev &R, uint32_t g) {
    const uint32_t r = R.ck_read[g];
    const uint32_t n = R.n[r];
    const uint32_t o0 = (g - R.ck_off[r]) * 32;
    if (o0 < n) {
        const uint32_t shift = R.shift[r], pos = R.pos[r];
        OpCur cur;
        op_at_block(R, r, g, shift + o0, cur);
        uint32_t tp, dl;
        col_tpos(R, r, pos, cur, shift + o0, tp, dl);
        R.ck_tpos[g] = tp;
        R.ck_delta[g] = (uint16_t)dl;
    }
    pack_block_slow(R, g);
}
// SEQ_NUM code of one BAM nibble ("=ACMGRSVTWYHKDBN" -> 4 0 1 6 2 4 4 4 3 4 4 4 4 4 4 5), as map_codes4 maps four
__device__ __forceinline__ uint32_t code_of_nibble(uint32_t v) { return (uint32_t)(0x5444444344426104ULL >> (4 * v)) & 15u; }
// The "holds something else than reference 3-mers" bit of a PLAIN block (32 M/=/X columns of one op), from what the pack
// thread has in registers: its 32 codes (memory-order words w0, w1) against the packed reference at tpos, and the two
// columns in front of the block, which are the two query bases before q0 when the op started at least two columns
// earlier (otherwise the block is simply called odd: K2 then walks it and finds nothing, which is exact, only slower).
// Same verdict as block_all_reference_at on the packed columns wherever it answers "all reference".
__device__ __forceinline__ bool plain_block_odd(uint64_t w0, uint64_t w1, uint32_t tpos, uint32_t o0, bool prev_in_op,
                                                const uint8_t *__restrict__ seq4, uint32_t q0,
                                                const uint32_t *__restrict__ refpk) {
    if (o0 == 0 || tpos < 2 || !prev_in_op) return true;
    if (((w0 | w1) & 0xCCCCCCCCCCCCCCCCULL) != 0) return true;
    const uint32_t k = tpos >> 3, sh = (tpos & 7) * 4;
    const uint32_t r0 = refpk[k], r1 = refpk[k + 1], r2 = refpk[k + 2], r3 = refpk[k + 3], r4 = refpk[k + 4];
    const uint32_t rp = k ? refpk[k - 1] : 0;
    const uint32_t w[4] = {(uint32_t)w0, (uint32_t)(w0 >> 32), (uint32_t)w1, (uint32_t)(w1 >> 32)};
    uint32_t diff = 0;
    diff |= __byte_perm(w[0], 0, 0x0123) ^ __funnelshift_l(r1, r0, sh);
    diff |= __byte_perm(w[1], 0, 0x0123) ^ __funnelshift_l(r2, r1, sh);
    diff |= __byte_perm(w[2], 0, 0x0123) ^ __funnelshift_l(r3, r2, sh);
    diff |= __byte_perm(w[3], 0, 0x0123) ^ __funnelshift_l(r4, r3, sh);
    if (diff) return true;
    // reference codes at tpos - 2, tpos - 1 (position i sits in nibble 7 - (i & 7) of word i >> 3)
    const uint32_t i2 = tpos - 2, i1 = tpos - 1;
    const uint32_t c2 = ((i2 >> 3) == k ? r0 : rp) >> (28 - 4 * (i2 & 7)) & 15u;
    const uint32_t c1 = ((i1 >> 3) == k ? r0 : rp) >> (28 - 4 * (i1 & 7)) & 15u;
    const uint32_t b2 = seq4[(q0 - 2) >> 1], b1 = seq4[(q0 - 1) >> 1];
    const uint32_t s2 = code_of_nibble(((q0 - 2) & 1) ? (b2 & 15u) : (b2 >> 4));
    const uint32_t s1 = code_of_nibble(((q0 - 1) & 1) ? (b1 & 15u) : (b1 >> 4));
    return s2 >= 4 || s1 >= 4 || s2 != c2 || s1 != c1;
}
template <int B>
__global__ void __launch_bounds__(kPackThreads, 4) k_pack_columns_batched(ReadsDev R, uint32_t n_blocks,
                                                                          const uint32_t *__restrict__ refpk,
                                                                          uint32_t *__restrict__ blk_odd) {
    __shared__ uint32_t q[kPackThreads * B], qn;
    if (threadIdx.x == 0) qn = 0;
    __syncthreads();
    uint32_t g[B], r[B], v[B];
#pragma unroll
    for (int u = 0; u < B; u++) {
        g[u] = (blockIdx.x * B + u) * kPackThreads + threadIdx.x;
        const uint32_t gc = min(g[u], n_blocks - 1);  // n_blocks > 0 (host wrapper)
        r[u] = R.ck_read[gc];
        v[u] = R.blk_op[gc];
    }
    uint32_t n[B], o0[B], shift[B], pos[B], opi[B];
    uint64_t soff[B], noff[B];
#pragma unroll
    for (int u = 0; u < B; u++) {
        n[u] = R.n[r[u]];
        o0[u] = (min(g[u], n_blocks - 1) - R.ck_off[r[u]]) * 32;
        shift[u] = R.shift[r[u]];
        pos[u] = R.pos[r[u]];
        opi[u] = R.op_off[r[u]] + v[u];
        soff[u] = R.seq_off[r[u]];
        noff[u] = R.nib_off[r[u]];
    }
    bool full[B];
    uint4 o[B];
#pragma unroll
    for (int u = 0; u < B; u++) {
        full[u] = g[u] < n_blocks && v[u] != 0xFFFFu && o0[u] + 32 <= n[u];  // 32 columns of the read, op known
        o[u] = make_uint4(0, 0, 0, 2);
        if (full[u]) o[u] = __ldg(R.ops + opi[u]);
    }
    bool plain[B];
    uint64_t hi[B], lo[B];
#pragma unroll
    for (int u = 0; u < B; u++) {
        const uint32_t op = o[u].w & 15, c_end = o[u].x + (o[u].w >> 4), col = shift[u] + o0[u];
        plain[u] = full[u] && op != 1 && op != 2 && col + 32 <= c_end;
        hi[u] = lo[u] = 0;
        if (plain[u]) load_seq32(R.blob + soff[u], o[u].y + (col - o[u].x), hi[u], lo[u]);
    }
#pragma unroll
    for (int u = 0; u < B; u++) {
        bool odd = false;
        if (plain[u]) {
            const uint32_t col = shift[u] + o0[u];
            const uint32_t tp = pos[u] + o[u].z + (col - o[u].x);  // col_tpos of an M/=/X column
            R.ck_tpos[g[u]] = tp;
            R.ck_delta[g[u]] = 0;
            uint64_t *out = (uint64_t *)(R.nib + noff[u] + (o0[u] >> 1));
            const uint64_t w0 = map_codes16(hi[u]), w1 = map_codes16(lo[u]);
            out[0] = w0;
            out[1] = w1;
            odd = plain_block_odd(w0, w1, tp, o0[u], col >= o[u].x + 2, R.blob + soff[u], o[u].y + (col - o[u].x), refpk);
        } else if (g[u] < n_blocks && n[u] && o0[u] <= n[u]) {
            q[atomicAdd(&qn, 1u)] = g[u];
            odd = o0[u] < n[u];  // an op boundary, an indel or the end of the read inside the block: K2 walks it
        }
        // bit g of blk_odd (what k_block_flags computes from the packed columns): the 32 blocks of a warp are one word
        const uint32_t bits = __ballot_sync(0xFFFFFFFFu, odd);
        if ((threadIdx.x & 31) == 0 && (g[u] >> 5) < ((n_blocks + 31) >> 5)) blk_odd[g[u] >> 5] = bits;
    }
    __syncthreads();
    const uint32_t nq = qn;
    for (uint32_t i = threadIdx.x; i < nq; i += kPackThreads) pack_block_queued(R, q[i]);
}
// ---- CIGAR -> op records (fill_with_cigar's bookkeeping, main.rs:386-440).  The raw CIGAR words of a read arrive in
// front of its SEQ bytes (one span per read over the link: 4 bytes per op instead of a host-built 16-byte record); a
// warp turns them into the records the other kernels read with one load: {first column, first query base, contig
// offset, raw word} of every op that owns alignment columns (M, I, D, =, X with a non-zero length).  Column, query
// and contig positions are exclusive prefix sums over the ops (S advances the query only, H nothing); the host has
// already rejected records with any other op, so no check is repeated here.
__global__ void __launch_bounds__(128) k_cigar_ops(const uint8_t *__restrict__ blob, const uint64_t *__restrict__ seq_off,
                                                   const uint32_t *__restrict__ n_cig, const uint32_t *__restrict__ op_off,
                                                   uint4 *__restrict__ ops, uint32_t n_reads) {
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= n_reads) return;
    const uint32_t nc = n_cig[r];
    const uint8_t *cg = blob + seq_off[r] - 4ull * nc;  // BAM records are not aligned: neither are their CIGAR words
    const uint32_t mis = (uint32_t)((uintptr_t)cg & 3), sh = mis * 8;
    const uint32_t *wp = reinterpret_cast<const uint32_t *>(cg - mis);
    uint32_t col = 0, q = 0, t = 0, w = op_off[r];
    for (uint32_t base = 0; base < nc; base += 32) {
        const uint32_t i = base + lane;
        uint32_t c = 0xF;  // an op code nothing counts
        if (i < nc) c = __funnelshift_r(wp[i], mis ? wp[i + 1] : 0u, sh);
        const uint32_t l = c >> 4, bit = 1u << (c & 15);
        uint32_t dc = (bit & 0x187u) ? l : 0u;  // M I D = X
        uint32_t dq = (bit & 0x193u) ? l : 0u;  // M I S = X
        uint32_t dt = (bit & 0x185u) ? l : 0u;  // M D = X
        const bool emit = dc != 0;
        const uint32_t l0 = dc, q0 = dq, t0 = dt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t a = __shfl_up_sync(0xFFFFFFFFu, dc, d), b = __shfl_up_sync(0xFFFFFFFFu, dq, d),
                           e = __shfl_up_sync(0xFFFFFFFFu, dt, d);
            if (lane >= (uint32_t)d) {
                dc += a;
                dq += b;
                dt += e;
            }
        }
        const uint32_t em = __ballot_sync(0xFFFFFFFFu, emit);
        if (emit) ops[w + __popc(em & ((1u << lane) - 1))] = make_uint4(col + dc - l0, q + dq - q0, t + dt - t0, c);
        col += __shfl_sync(0xFFFFFFFFu, dc, 31);
        q += __shfl_sync(0xFFFFFFFFu, dq, 31);
        t += __shfl_sync(0xFFFFFFFFu, dt, 31);
        w += __popc(em);
    }
}
void cigar_ops(const uint8_t *d_blob, const uint64_t *d_seq_off, const uint32_t *d_n_cig, const uint32_t *d_op_off,
               uint4 *d_ops, uint32_t n_reads, cudaStream_t s) {
    if (n_reads)
        NP2_K(k_cigar_ops)<<<cdiv((uint64_t)n_reads * 32, 128), 128, 0, s>>>(d_blob, d_seq_off, d_n_cig, d_op_off, d_ops, n_reads);
}
void trim_scan(const ReadsDev &r, const uint8_t *d_ref, uint32_t L, cudaStream_t s) {
    if (r.n_reads) NP2_K(k_trim_scan)<<<cdiv((uint64_t)r.n_reads * 32, 128), 128, 0, s>>>(r, d_ref, L);
}
// returns true when the per-block "not all reference" bits (d_blk_odd) were produced along the way; the
// one-block-per-thread form (NP2_PACK_BATCH=1, A/B runs) leaves them to block_flags
bool pack_columns(const ReadsDev &r, const uint8_t *d_ref, uint32_t n_blocks, const uint32_t *d_refpk, uint32_t *d_blk_odd,
                  cudaStream_t s) {
    if (!r.n_reads || !n_blocks) return false;
    static const int batch = [] {
        const char *e = getenv("NP2_PACK_BATCH");
        return e ? atoi(e) : kPackBatchDefault;
    }();
    if (batch <= 1) {
        NP2_K(k_pack_columns)<<<cdiv(n_blocks, kPackThreads), kPackThreads, 0, s>>>(r, d_ref, n_blocks);
        return false;
    }
    NP2_K(k_pack_columns_batched<2>)<<<cdiv(n_blocks, 2 * kPackThreads), kPackThreads, 0, s>>>(r, n_blocks, d_refpk, d_blk_odd);
    return true;
}

/* =============================================================== K2: pileup */

__global__ void k_cover_diff(ReadsDev R, const uint8_t *__restrict__ blank, int32_t *diff) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r == R.n_reads) atomicAdd(&diff[0], 1);  // the ref read spans [0, L - 1] (main.rs:1732-1739)
    if (r >= R.n_reads || blank[r] || R.n[r] == 0) return;
    atomicAdd(&diff[R.t_s[r]], 1);
    atomicAdd(&diff[R.t_e[r] + 1], -1);
}
void cover_diff(const ReadsDev &r, const uint8_t *d_blank, int32_t *d_diff, cudaStream_t s) {
    NP2_K(k_cover_diff)<<<cdiv(r.n_reads + 1, kThreads), kThreads, 0, s>>>(r, d_blank, d_diff);
}
void cover_scan(int32_t *d_cover, uint32_t n, ScanPool &pool, cudaStream_t s) {
    scan_launch(ScanInclusiveI32{{}, d_cover}, nullptr, 0, n, pool, s);
}
__global__ void k_counts_init(CountsDev cd) {
    const uint32_t i = threadIdx.x;
    if (i < C_COUNT) cd.c[i] = 0u;
    if (i < Q_COUNT) cd.q[i] = 0;
}
void counts_init(CountsDev cd, cudaStream_t s) { NP2_K(k_counts_init)<<<1, 32, 0, s>>>(cd); }

__device__ __forceinline__ uint32_t nib_at(const uint8_t *__restrict__ nib, uint32_t o) {
    uint32_t b = nib[o >> 1];
    return (o & 1) ? (b & 15) : (b >> 4);
}
// delta of column o = number of consecutive insertion columns ending at o
__device__ __forceinline__ uint32_t delta_scanback(const uint8_t *__restrict__ nib, uint32_t o) {
    uint32_t d = 0;
    while ((nib_at(nib, o) & 8) && o > 0) {
        d++;
        o--;
    }
    return d;
}

// Fast check (the overwhelmingly common block): 32 plain columns (codes 0-3, no insertion, no gap) that equal the
// reference, preceded by two plain matching columns => every 3-mer of the block is the reference 3-mer and nothing
// has to be emitted.  Four word compares instead of 32 column steps.  Also true for blocks that hold no columns.
__device__ __forceinline__ bool block_all_reference(const ReadsDev &R, uint32_t g, const uint8_t *__restrict__ blank,
                                                    const uint8_t *__restrict__ code,
                                                    const uint32_t *__restrict__ refpk) {
    const uint32_t r = R.ck_read[g];
    if (blank[r]) return true;
    const uint32_t n = R.n[r];
    const uint32_t o0 = (g - R.ck_off[r]) * 32;
    if (o0 >= n) return true;
    if (o0 == 0 || o0 + 32 > n) return false;
    const uint8_t *nib = R.nib + R.nib_off[r];
    const uint4 w4 = *reinterpret_cast<const uint4 *>(nib + (o0 >> 1));
    if (((w4.x | w4.y | w4.z | w4.w) & 0xCCCCCCCCu) != 0) return false;
    const uint32_t tpos = R.ck_tpos[g];
    const uint32_t pb = nib[(o0 >> 1) - 1];
    if ((pb & 0xCCu) != 0 || (pb >> 4) != code[tpos - 2] || (pb & 15u) != code[tpos - 1]) return false;
    const uint32_t k = tpos >> 3, sh = (tpos & 7) * 4;
    const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
    uint32_t diff = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t rj = __byte_perm(w[j], 0, 0x0123);  // column 8j in the most significant nibble
        diff |= rj ^ __funnelshift_l(refpk[k + j + 1], refpk[k + j], sh);
    }
    return diff == 0;
}

// Walks the 32 columns of global block g and calls f(p, bases, delta1) for every 3-mer that is NOT the
// reference 3-mer of its position (those are counted as cover[p] - #others, see k_pileup_stripe).
template <class F>
__device__ __forceinline__ void scan_block32(const ReadsDev &R, uint32_t g, const uint8_t *__restrict__ code, F f) {
    const uint32_t r = R.ck_read[g];
    const uint32_t n = R.n[r];
    const uint32_t o0 = (g - R.ck_off[r]) * 32;
    const uint8_t *nib = R.nib + R.nib_off[r];
    const uint4 w4 = *reinterpret_cast<const uint4 *>(nib + (o0 >> 1));
    const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
    uint32_t tpos = R.ck_tpos[g], delta = R.ck_delta[g];
    uint32_t q1 = 15, q2 = 15, ins1 = 0, ins2 = 0, d1 = 0, d2 = 1;  // heads (main.rs:579-580)
    if (o0 > 0) {
        uint32_t pb = nib[(o0 >> 1) - 1];
        uint32_t v2 = pb & 15, v1 = pb >> 4;  // columns o0-1, o0-2
        q2 = v2 & 7;
        ins2 = v2 >> 3;
        q1 = v1 & 7;
        ins1 = v1 >> 3;
        uint32_t first = (w[0] >> 4) & 15;  // column o0
        if (first & 8) d2 = delta - 1;
        else d2 = ins2 ? delta_scanback(nib, o0 - 1) : 0;
        if (ins2) d1 = d2 - 1;
        else d1 = ins1 ? delta_scanback(nib, o0 - 2) : 0;
    }
    uint32_t rc0 = code[tpos], rc1 = tpos >= 1 ? code[tpos - 1] : 255u, rc2 = tpos >= 2 ? code[tpos - 2] : 255u;
    const uint32_t cnt = min(32u, n - o0);
#pragma unroll
    for (uint32_t i = 0; i < 32; i++) {
        if (i >= cnt) break;
        const uint32_t byte = (w[i >> 3] >> (8 * ((i >> 1) & 3))) & 255;
        const uint32_t v = (i & 1) ? (byte & 15) : (byte >> 4);
        const uint32_t q3 = v & 7, ins3 = v >> 3;
        if (i > 0) {
            if (ins3) delta++;
            else {
                delta = 0;
                tpos++;
                rc2 = rc1;
                rc1 = rc0;
                rc0 = code[tpos];
            }
        }
        const uint32_t o = o0 + i;
        uint32_t bases, dl1;
        bool dense = false;
        if (o == 0) {
            bases = 0x4000u | 15u << 8 | 15u << 4 | q3;
            dl1 = 0;
        } else if (o == 1) {
            bases = (ins3 ? 0x1000u : 0u) | 15u << 8 | q2 << 4 | q3;
            dl1 = 1;
        } else {
            bases = (ins2 ? 0x4000u : 0u) | (ins3 ? 0x1000u : 0u) | q1 << 8 | q2 << 4 | q3;
            dl1 = d1;
            dense = !(ins1 | ins2 | ins3) && q1 == rc2 && q2 == rc1 && q3 == rc0;
        }
        if (!dense) f(tpos, bases, dl1 & 0xFFFFu);
        q1 = q2;
        ins1 = ins2;
        d1 = d2;
        q2 = q3;
        ins2 = ins3;
        d2 = delta;
    }
}

// block_all_reference for a block whose read is already known: 32 plain columns o0 .. o0 + 31 of a read of n columns
// (nib = its packed columns, tpos = t_pos of column o0) that equal the reference and follow two plain matching columns
__device__ __forceinline__ bool block_all_reference_at(const uint8_t *__restrict__ nib, uint32_t n, uint32_t o0,
                                                       uint32_t tpos, const uint8_t *__restrict__ code,
                                                       const uint32_t *__restrict__ refpk) {
    if (o0 >= n) return true;
    if (o0 == 0 || o0 + 32 > n || tpos < 2) return false;
    const uint4 w4 = *reinterpret_cast<const uint4 *>(nib + (o0 >> 1));
    if (((w4.x | w4.y | w4.z | w4.w) & 0xCCCCCCCCu) != 0) return false;
    const uint32_t pb = nib[(o0 >> 1) - 1];
    if ((pb & 0xCCu) != 0 || (pb >> 4) != code[tpos - 2] || (pb & 15u) != code[tpos - 1]) return false;
    const uint32_t k = tpos >> 3, sh = (tpos & 7) * 4;
    const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
    uint32_t diff = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t rj = __byte_perm(w[j], 0, 0x0123);  // column 8j in the most significant nibble
        diff |= rj ^ __funnelshift_l(refpk[k + j + 1], refpk[k + j], sh);
    }
    return diff == 0;
}

/* ---------------------------------------------------------------------------------------------------------------
 * K2 proper: update_msas + Msa::push / sort / coverage (main.rs:576-589, 193-241) for one STRIPE of kStripeW contig
 * positions per CTA, entirely in shared memory.
 *
 * Reads are coordinate-sorted, so the reads that can cover a stripe sit in a window of the read array; for each of
 * them the checkpoints give the 32-column blocks that can hold a column of the stripe.  Those (read, block) items are
 * tested against the reference with four word compares (block_all_reference); the few odd ones are walked column by
 * column (scan_block32) and their non-reference 3-mers land in a shared-memory record list.  Which blocks those are
 * does not depend on the iteration: k_block_flags + k_stripe_odd list them per stripe once per job, this kernel only
 * reads its list (and skips the reads the iteration has blanked) — finding them here, 64 threads per batch of reads
 * with the other 192 at a barrier, was 45 % of the kernel's time (profiles/r02u_source_hotspots.txt).  The list is bucketed by
 * position (counting sort), identical 3-mers of a position are merged into Msa entries (count, first read), the
 * entries of every position are put into Msa::sort order (b3.delta, then first read — unique per entry, so the result
 * does not depend on the order the records were found in), and per position the kernel writes what the DP needs:
 * where its entries are, how many, the count of the implicit reference 3-mer (cover - sum of the others with
 * b3.delta == 0), whether it has more than one entry, and how many consensus bases a single-entry position emits.
 * Nothing but the finished entries ever goes to HBM: no record buffer, no global sort, no separate group / finalize
 * passes.  Entries of a stripe get their place in the global arrays by ONE atomic reservation (their order across
 * stripes is irrelevant: every position knows its own offset and count).
 *
 * A stripe whose records do not fit the list (kStripeRmax) is split in halves by position, recursively; the walk is
 * repeated for each half.  WRITE = false only counts entries (exact mode sizes the arrays from that). */
constexpr int kStripeThreads = 256;
constexpr int kStripeRmax = 1024;   // records of one position range held in shared memory
template <int kStripeW>
struct StripeSmem {
    uint32_t bd[kStripeRmax], rd[kStripeRmax], first[kStripeRmax];
    uint16_t kp[kStripeRmax], perm[kStripeRmax], cnt[kStripeRmax], slot[kStripeRmax];
    uint32_t off[kStripeW + 2], goff[kStripeW + 2];
    uint32_t stk_a[12], stk_b[12];
    // the stripe's window of the per-position inputs (coverage, reference codes), brought in by TMA bulk copies while
    // the reads are walked.  (Staging the per-position outputs too and writing them back with bulk stores was measured:
    // the extra 15 KB of shared memory cost more occupancy than the stores cost LSU time, 0.35 vs 0.30 ms.)
    alignas(16) int32_t w_cover[kStripeW];
    alignas(16) uint8_t w_code[kStripeW];
    alignas(8) uint64_t bar;
    uint32_t nrec, nodd, base, ok;
    unsigned long long score;  // 10 * count - 4 * coverage over the single-entry positions of the range (main.rs:1659)
    int sp;
};
// positions per stripe: 1024 (NP2_STRIPE_W=512 for A/B runs)
static int stripe_w() {
    static const int w = [] {
        const char *e = getenv("NP2_STRIPE_W");
        return e && atoi(e) == 512 ? 512 : 1024;
    }();
    return w;
}
uint32_t pileup_stripes(uint32_t L) { return cdiv(L, stripe_w()); }
uint32_t pileup_stripe_width() { return (uint32_t)stripe_w(); }

template <int kStripeW, bool WRITE>
__global__ void __launch_bounds__(kStripeThreads) k_pileup_stripe(ReadsDev R, const uint8_t *__restrict__ blank,
                                                                  const uint8_t *__restrict__ code,
                                                                  const uint32_t *__restrict__ odd_off,
                                                                  const uint32_t *__restrict__ odd_list, MsaDev m,
                                                                  uint32_t cap_g, CountsDev cd,
                                                                  uint32_t *__restrict__ n_emit) {
    extern __shared__ __align__(16) unsigned char stripe_smem_raw[];
    StripeSmem<kStripeW> &S = *reinterpret_cast<StripeSmem<kStripeW> *>(stripe_smem_raw);
    constexpr int kPer = kStripeW / kStripeThreads;  // positions per thread in the block scans
    typedef cub::BlockScan<uint32_t, kStripeThreads> BS;
    __shared__ typename BS::TempStorage bs_tmp;
    namespace ptx = cuda::ptx;
    const uint32_t tid = threadIdx.x, L = m.L;
    if (cd.c[C_ABORT]) return;
    if (tid == 0) {
        S.sp = 1;
        S.stk_a[0] = blockIdx.x * kStripeW;
        S.stk_b[0] = min(blockIdx.x * kStripeW + kStripeW, L);
        ptx::mbarrier_init(&S.bar, 1);
        ptx::fence_proxy_async(ptx::space_shared);  // barrier initialisation visible to the async proxy
    }
    uint32_t bar_phase = 0;
    for (;;) {
        __syncthreads();
        if (S.sp == 0) break;
        const uint32_t a = S.stk_a[S.sp - 1], b = S.stk_b[S.sp - 1];
        __syncthreads();
        if (tid == 0) {
            S.sp--;
            S.nrec = 0;
            S.score = 0;
        }
        __syncthreads();
        // TMA: coverage and reference codes of [a, b) into shared memory, in flight while the reads are walked.  Whole
        // 16-byte units only (stripes start at multiples of 1024; a split range falls back to plain loads).
        const uint32_t Wr = b - a;
        const bool tma_in = WRITE && (a & 15) == 0 && (Wr & 15) == 0;
        if (tma_in && tid == 0) {
            ptx::mbarrier_arrive_expect_tx(ptx::sem_release, ptx::scope_cta, ptx::space_shared, &S.bar, Wr * 5);
            ptx::cp_async_bulk(ptx::space_cluster, ptx::space_global, S.w_cover, m.cover + a, Wr * 4, &S.bar);
            ptx::cp_async_bulk(ptx::space_cluster, ptx::space_global, S.w_code, m.code + a, Wr, &S.bar);
        }
        auto append = [&](uint32_t p, uint32_t bases, uint32_t dl1, uint32_t order) {
            if (p < a || p >= b) return;
            const uint32_t i = atomicAdd(&S.nrec, 1u);
            if (i < kStripeRmax) {
                S.kp[i] = (uint16_t)(p - a);
                S.bd[i] = bases << 16 | dl1;
                S.rd[i] = order;
            }
        };
        // the reference read's two head 3-mers (main.rs:1732-1739, 579-584)
        if (tid == 0) {
            append(0, 0x4000u | 15u << 8 | 15u << 4 | code[0], 0, 0);
            append(1, 15u << 8 | (uint32_t)code[0] << 4 | code[1], 1, 0);
        }
        // The blocks that can hold a column of this stripe and are not all reference were listed once per job
        // (k_stripe_odd_*): walk them, skipping the reads that are blank in this iteration.  (A split range walks the
        // whole stripe's list again; append() keeps what falls into [a, b).)
        {
            const uint32_t l0 = odd_off[blockIdx.x], l1 = odd_off[blockIdx.x + 1];
            for (uint32_t t = l0 + tid; t < l1; t += kStripeThreads) {
                const uint32_t g = odd_list[t];
                const uint32_t r = R.ck_read[g];
                if (blank[r]) continue;
                scan_block32(R, g, code, [&](uint32_t p, uint32_t bases, uint32_t dl1) { append(p, bases, dl1, r + 1); });
            }
        }
        __syncthreads();
        const uint32_t nrec = S.nrec;
        if (nrec > kStripeRmax) {  // split by position and walk again
            if (tma_in) {  // the window is not used: let the copies land before the buffers are reused
                while (!ptx::mbarrier_try_wait_parity(&S.bar, bar_phase)) {
                }
                bar_phase ^= 1;
            }
            if (tid == 0) {
                if (b - a <= 1) {
                    atomicExch(cd.c + C_PERR, 5u);  // one position with more records than the list holds
                } else {
                    const uint32_t mid = a + (b - a) / 2;
                    S.stk_a[S.sp] = mid;
                    S.stk_b[S.sp] = b;
                    S.stk_a[S.sp + 1] = a;
                    S.stk_b[S.sp + 1] = mid;
                    S.sp += 2;
                }
            }
            continue;
        }
        const uint32_t W = b - a;
        // ---- counting sort by position
        for (uint32_t i = tid; i < kStripeW + 2; i += kStripeThreads) S.off[i] = 0;
        __syncthreads();
        for (uint32_t i = tid; i < nrec; i += kStripeThreads) atomicAdd(&S.off[S.kp[i]], 1u);
        __syncthreads();
        {
            uint32_t v[kPer], sum = 0, ex;
#pragma unroll
            for (int u = 0; u < kPer; u++) sum += v[u] = S.off[kPer * tid + u];
            BS(bs_tmp).ExclusiveSum(sum, ex);
#pragma unroll
            for (int u = 0; u < kPer; u++) {
                S.off[kPer * tid + u] = ex;
                S.goff[kPer * tid + u] = ex;  // cursors
                ex += v[u];
            }
            if (tid == kStripeThreads - 1) S.off[kStripeW] = ex;
        }
        __syncthreads();
        for (uint32_t i = tid; i < nrec; i += kStripeThreads) S.perm[atomicAdd(&S.goff[S.kp[i]], 1u)] = (uint16_t)i;
        __syncthreads();
        for (uint32_t i = tid; i < kStripeW + 2; i += kStripeThreads) S.goff[i] = 0;
        __syncthreads();
        // ---- entries: the first record of every distinct 3-mer of a position counts its copies
        for (uint32_t j = tid; j < nrec; j += kStripeThreads) {
            const uint32_t i = S.perm[j], k = S.kp[i], bdv = S.bd[i];
            const uint32_t lo = S.off[k], hi = S.off[k + 1];
            bool head = true;
            for (uint32_t x = lo; x < j; x++)
                if (S.bd[S.perm[x]] == bdv) {
                    head = false;
                    break;
                }
            uint32_t slot = 0xFFFFu;
            if (head) {
                uint32_t c = 1, f = S.rd[i];
                for (uint32_t x = j + 1; x < hi; x++) {
                    const uint32_t i2 = S.perm[x];
                    if (S.bd[i2] == bdv) {
                        c++;
                        f = min(f, S.rd[i2]);
                    }
                }
                slot = atomicAdd(&S.goff[k], 1u);
                if (c > 0xFFFFu) atomicExch(cd.c + C_PERR, 6u);
                S.cnt[j] = (uint16_t)min(c, 0xFFFFu);
                S.first[j] = f;
            }
            S.slot[j] = (uint16_t)slot;
        }
        __syncthreads();
        uint32_t ng_total;
        {
            uint32_t v[kPer], sum = 0, ex;
#pragma unroll
            for (int u = 0; u < kPer; u++) sum += v[u] = S.goff[kPer * tid + u];
            BS(bs_tmp).ExclusiveSum(sum, ex, ng_total);
            __syncthreads();
#pragma unroll
            for (int u = 0; u < kPer; u++) {
                S.goff[kPer * tid + u] = ex;
                ex += v[u];
            }
            if (tid == kStripeThreads - 1) S.goff[kStripeW] = ex;
        }
        if (tid == 0) {
            const uint32_t base = atomicAdd(cd.c + C_G, ng_total);
            atomicAdd(cd.c + C_NREC, nrec);
            S.base = base;
            S.ok = 1;
            if (WRITE && (uint64_t)base + ng_total > cap_g) {
                atomicExch(cd.c + C_ABORT, 1u);
                S.ok = 0;
            }
        }
        __syncthreads();
        if (!WRITE || !S.ok) continue;
        const uint32_t base = S.base;
        for (uint32_t j = tid; j < nrec; j += kStripeThreads) {
            const uint32_t slot = S.slot[j];
            if (slot == 0xFFFFu) continue;
            const uint32_t i = S.perm[j];
            const uint32_t g = base + S.goff[S.kp[i]] + slot;
            m.g_bases[g] = (uint16_t)(S.bd[i] >> 16);
            m.g_delta[g] = (uint16_t)S.bd[i];
            m.g_count[g] = S.cnt[j];
            m.g_first[g] = S.first[j];
        }
        __syncthreads();
        // ---- per position: Msa::sort order (main.rs:193-229), reference 3-mer count, articulation flag
        if (tma_in) {
            while (!ptx::mbarrier_try_wait_parity(&S.bar, bar_phase)) {
            }
            bar_phase ^= 1;
        }
        long long score = 0;
        for (uint32_t k = tid; k < W; k += kStripeThreads) {
            const uint32_t p = a + k;
            const uint32_t lo = base + S.goff[k], hi = base + S.goff[k + 1];
            uint32_t sum0 = 0;
            for (uint32_t i = lo; i < hi; i++) {
                uint16_t bs = m.g_bases[i], dl = m.g_delta[i];
                uint32_t cn = m.g_count[i], fr = m.g_first[i];
                uint32_t kd = kmer_b3delta(bs, dl);
                if (kd == 0) sum0 += cn;
                uint32_t j = i;
                while (j > lo) {
                    uint32_t kd2 = kmer_b3delta(m.g_bases[j - 1], m.g_delta[j - 1]);
                    if (kd2 < kd || (kd2 == kd && m.g_first[j - 1] <= fr)) break;
                    m.g_bases[j] = m.g_bases[j - 1];
                    m.g_delta[j] = m.g_delta[j - 1];
                    m.g_count[j] = m.g_count[j - 1];
                    m.g_first[j] = m.g_first[j - 1];
                    j--;
                }
                if (j != i) {
                    m.g_bases[j] = bs;
                    m.g_delta[j] = dl;
                    m.g_count[j] = cn;
                    m.g_first[j] = fr;
                }
            }
            if (hi - lo > 0xFFFFu) atomicExch(cd.c + C_PERR, 6u);
            const uint32_t cov = tma_in ? (uint32_t)S.w_cover[k] : (uint32_t)m.cover[p];
            const uint32_t cde = tma_in ? S.w_code[k] : m.code[p];
            const bool multi = p < 2 || hi > lo;
            const uint32_t dense = p >= 2 ? cov - sum0 : 0, ne = multi ? 0 : (cde != 4);
            m.sp_off[p] = lo;
            m.sp_cnt[p] = (uint16_t)min(hi - lo, 0xFFFFu);
            m.dense_cnt[p] = dense;
            m.multi[p] = multi;
            n_emit[p] = ne;
            if (!multi) score += 6ll * cov;  // its only entry is the reference 3-mer: count == coverage
        }
        for (int d = 16; d > 0; d >>= 1) score += __shfl_xor_sync(0xFFFFFFFFu, score, d);
        if ((tid & 31) == 0 && score) atomicAdd(&S.score, (unsigned long long)score);
        __syncthreads();
        if (tid == 0 && S.score) atomicAdd(cd.q + Q_TOTAL, S.score);
    }
}
// bit g of blk_odd = 32-column block g holds something else than reference 3-mers (or cannot be decided by the word
// test: first / last block of a read).  A property of the read and the contig only: computed once per job.
__global__ void __launch_bounds__(kThreads) k_block_flags(ReadsDev R, uint32_t n_blocks, const uint8_t *__restrict__ code,
                                                          const uint32_t *__restrict__ refpk, uint32_t *__restrict__ blk_odd) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    bool odd = false;
    if (g < n_blocks) {
        const uint32_t r = R.ck_read[g];
        const uint32_t n = R.n[r], o0 = (g - R.ck_off[r]) * 32;
        if (o0 < n) odd = !block_all_reference_at(R.nib + R.nib_off[r], n, o0, R.ck_tpos[g], code, refpk);
    }
    const uint32_t bits = __ballot_sync(0xFFFFFFFFu, odd);
    if ((threadIdx.x & 31) == 0) blk_odd[g >> 5] = bits;
}
void block_flags(const ReadsDev &r, uint32_t n_blocks, const uint8_t *d_code, const uint32_t *d_refpk, uint32_t *d_blk_odd,
                 cudaStream_t s) {
    if (n_blocks) NP2_K(k_block_flags)<<<cdiv(n_blocks, kThreads), kThreads, 0, s>>>(r, n_blocks, d_code, d_refpk, d_blk_odd);
}
// first_ge[i] = first read whose record position is >= i * W (i = 0 .. stripes): the read window of a stripe
__global__ void k_stripe_reads(const uint32_t *__restrict__ pos, uint32_t n_reads, uint32_t n_entries, uint32_t W,
                               uint32_t *__restrict__ first_ge) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_entries) return;
    const uint64_t want = (uint64_t)i * W;
    uint32_t lo = 0, hi = n_reads;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (pos[mid] < want) lo = mid + 1;
        else hi = mid;
    }
    first_ge[i] = lo;
}
void stripe_reads(const ReadsDev &r, uint32_t L, uint32_t *d_first_ge, cudaStream_t s) {
    const uint32_t n = pileup_stripes(L) + 1;
    NP2_K(k_stripe_reads)<<<cdiv(n, kThreads), kThreads, 0, s>>>(r.pos, r.n_reads, n, (uint32_t)stripe_w(), d_first_ge);
}
// ---- per-stripe lists of the not-all-reference blocks, once per job.  For every stripe: the reads whose record
// position lies in [stripe start - max_span, stripe end) (per-stripe read index), for each of them the 32-column blocks
// that can hold a column of the stripe — from the last block that starts before the stripe to the last block that
// starts before its end.  HiFi alignments have few indels, so the range is guessed from the distance to the read's
// start with two blocks of margin (a block outside the stripe costs a walk that emits nothing) and VERIFIED with two
// checkpoints; a long indel makes the guess useless, then it is a binary search.  Of those blocks the ones whose
// k_block_flags bit is set go on the stripe's list.  Pass 0 counts, pass 1 (after the offsets scan) writes.  Neither
// the trim nor the flags depend on which reads an iteration blanks: K2 filters those while it walks.
__device__ __forceinline__ void stripe_block_range(const ReadsDev &R, uint32_t r, uint32_t a, uint32_t b, uint32_t &g_lo,
                                                   uint32_t &g_hi) {
    g_lo = g_hi = 0;
    const uint32_t n = R.n[r], ts = R.t_s[r], te = R.t_e[r], c0 = R.ck_off[r];
    if (!n || te < a || ts >= b) return;
    const uint32_t nblk = (n + 31) >> 5;
    const uint32_t *ck = R.ck_tpos + c0;
    const uint32_t ga = a > ts ? (a - ts) >> 5 : 0;
    uint32_t b0 = ga > 2 ? min(ga - 2, nblk - 1) : 0, e = min(nblk, ga + ((b - a) >> 5) + 4);
    const uint32_t ck_b0 = ck[b0], ck_e = e < nblk ? ck[e] : 0xFFFFFFFFu;
    if ((b0 > 0 && ck_b0 >= a) || ck_e < b) {  // exact range
        uint32_t lo = 0, hi = nblk;  // blocks whose first t_pos is < a
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (ck[mid] < a) lo = mid + 1;
            else hi = mid;
        }
        b0 = lo ? lo - 1 : 0;
        hi = nblk;  // blocks whose first t_pos is < b
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (ck[mid] < b) lo = mid + 1;
            else hi = mid;
        }
        e = max(lo, b0);
    }
    g_lo = c0 + b0;
    g_hi = c0 + e;
}
constexpr int kOddThreads = 64;
template <bool WRITE>
__global__ void __launch_bounds__(kOddThreads) k_stripe_odd(ReadsDev R, const uint32_t *__restrict__ blk_odd,
                                                            const uint32_t *__restrict__ first_ge, uint32_t L, uint32_t W,
                                                            uint32_t max_span, uint32_t *__restrict__ odd_cnt,
                                                            const uint32_t *__restrict__ odd_off,
                                                            uint32_t *__restrict__ odd_list, uint32_t cap) {
    __shared__ uint32_t s_warp[kOddThreads / 32], s_base;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t a = blockIdx.x * W, b = min(a + W, L);
    const uint32_t lo_pos = a > max_span ? a - max_span : 0;
    const uint32_t r_lo = first_ge[lo_pos / W], r_hi = first_ge[min((b + W - 1) / W, gridDim.x)];
    uint32_t total = 0;  // WRITE: entries of this stripe written by the batches before
    if (WRITE && tid == 0) s_base = odd_off[blockIdx.x];
    for (uint32_t rb = r_lo; rb < r_hi; rb += kOddThreads) {
        const uint32_t r = rb + tid;
        uint32_t g_lo = 0, g_hi = 0, n_odd = 0;
        if (r < r_hi) {
            stripe_block_range(R, r, a, b, g_lo, g_hi);
            for (uint32_t w = g_lo >> 5; g_hi > g_lo && w <= (g_hi - 1) >> 5; w++) {
                uint32_t bits = blk_odd[w];
                if (w == g_lo >> 5) bits &= 0xFFFFFFFFu << (g_lo & 31);
                if (w == (g_hi - 1) >> 5) bits &= 0xFFFFFFFFu >> (31 - ((g_hi - 1) & 31));
                n_odd += __popc(bits);
            }
        }
        if (!WRITE) {
            total += n_odd;
            continue;
        }
        // exclusive offsets inside the batch
        uint32_t incl = n_odd;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= (uint32_t)d) incl += t;
        }
        __syncthreads();  // s_warp of the batch before has been read (and s_base is visible)
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t ex = incl - n_odd, all = 0;
#pragma unroll
        for (int w = 0; w < kOddThreads / 32; w++) {
            if (w < (int)warp) ex += s_warp[w];
            all += s_warp[w];
        }
        uint32_t w_out = s_base + total + ex;
        for (uint32_t w = g_lo >> 5; n_odd && w <= (g_hi - 1) >> 5; w++) {
            uint32_t bits = blk_odd[w];
            if (w == g_lo >> 5) bits &= 0xFFFFFFFFu << (g_lo & 31);
            if (w == (g_hi - 1) >> 5) bits &= 0xFFFFFFFFu >> (31 - ((g_hi - 1) & 31));
            for (; bits; bits &= bits - 1, w_out++)
                if (w_out < cap) odd_list[w_out] = (w << 5) + (uint32_t)__ffs(bits) - 1;
        }
        total += all;
    }
    if (!WRITE) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) total += __shfl_xor_sync(0xFFFFFFFFu, total, d);
        if (lane == 0) s_warp[warp] = total;
        __syncthreads();
        if (tid == 0) {
            uint32_t t = 0;
            for (int w = 0; w < kOddThreads / 32; w++) t += s_warp[w];
            odd_cnt[blockIdx.x] = t;
        }
    }
}
void stripe_odd_count(const ReadsDev &r, const uint32_t *d_blk_odd, const uint32_t *d_first_ge, uint32_t L, uint32_t max_span,
                      uint32_t *d_odd_cnt, cudaStream_t s) {
    NP2_K(k_stripe_odd<false>)<<<pileup_stripes(L), kOddThreads, 0, s>>>(r, d_blk_odd, d_first_ge, L, (uint32_t)stripe_w(), max_span,
                                                                       d_odd_cnt, nullptr, nullptr, 0);
}
// d_odd_off[stripes + 1] = exclusive sum of the counts; the total also goes to *d_total
void stripe_odd_offsets(const uint32_t *d_odd_cnt, uint32_t *d_odd_off, uint32_t L, uint32_t *d_total, ScanPool &pool,
                        cudaStream_t s) {
    ScanOffsets<uint32_t, uint32_t> f;
    f.in = d_odd_cnt;
    f.out = d_odd_off;
    f.c_slot = d_total;
    f.q_slot = nullptr;
    f.cap = ~0ULL;
    f.abort = nullptr;
    scan_launch(f, nullptr, 0, pileup_stripes(L), pool, s);
}
void stripe_odd_fill(const ReadsDev &r, const uint32_t *d_blk_odd, const uint32_t *d_first_ge, uint32_t L, uint32_t max_span,
                     const uint32_t *d_odd_off, uint32_t *d_odd_list, uint32_t cap, cudaStream_t s) {
    NP2_K(k_stripe_odd<true>)<<<pileup_stripes(L), kOddThreads, 0, s>>>(r, d_blk_odd, d_first_ge, L, (uint32_t)stripe_w(), max_span,
                                                                      nullptr, d_odd_off, d_odd_list, cap);
}
template <int W>
static void pileup_stripe_w(const ReadsDev &r, const uint8_t *d_blank, const uint8_t *d_code, const uint32_t *d_odd_off,
                            const uint32_t *d_odd_list, MsaDev m, uint32_t cap_g, CountsDev cd, uint32_t *d_n_emit,
                            bool count_only, cudaStream_t s) {
    static bool attr_done = false;
    const int smem = (int)sizeof(StripeSmem<W>);
    if (!attr_done) {
        NP2_CUDA(cudaFuncSetAttribute(k_pileup_stripe<W, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        NP2_CUDA(cudaFuncSetAttribute(k_pileup_stripe<W, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_done = true;
    }
    const uint32_t grid = cdiv(m.L, W);
    if (count_only)
        NP2_K((k_pileup_stripe<W, false>))<<<grid, kStripeThreads, smem, s>>>(r, d_blank, d_code, d_odd_off, d_odd_list, m, cap_g, cd,
                                                                           d_n_emit);
    else
        NP2_K((k_pileup_stripe<W, true>))<<<grid, kStripeThreads, smem, s>>>(r, d_blank, d_code, d_odd_off, d_odd_list, m, cap_g, cd,
                                                                          d_n_emit);
}
// count_only: only C_G / C_NREC are produced (exact mode sizes the entry arrays from them)
void pileup_stripe(const ReadsDev &r, const uint8_t *d_blank, const uint8_t *d_code, const uint32_t *d_odd_off,
                   const uint32_t *d_odd_list, MsaDev m, uint32_t cap_g, CountsDev cd, uint32_t *d_n_emit, bool count_only,
                   cudaStream_t s) {
    if (stripe_w() == 512)
        pileup_stripe_w<512>(r, d_blank, d_code, d_odd_off, d_odd_list, m, cap_g, cd, d_n_emit, count_only, s);
    else
        pileup_stripe_w<1024>(r, d_blank, d_code, d_odd_off, d_odd_list, m, cap_g, cd, d_n_emit, count_only, s);
}
__global__ void k_counts_reset_pileup(CountsDev cd) {
    cd.c[C_G] = 0;
    cd.c[C_NREC] = 0;
}
void counts_reset_pileup(CountsDev cd, cudaStream_t s) { NP2_K(k_counts_reset_pileup)<<<1, 1, 0, s>>>(cd); }

struct PredRunStart {
    const uint8_t *multi;
    __device__ bool operator()(uint32_t p) const { return multi[p] && (p == 0 || !multi[p - 1]); }
};
void runs_select(const uint8_t *d_multi, uint32_t L, uint32_t *d_run_start, uint32_t cap_runs, CountsDev cd, ScanPool &pool,
                 cudaStream_t s) {
    ScanSelect<PredRunStart> f;
    f.pred = PredRunStart{d_multi};
    f.out = d_run_start;
    f.count = cd.c + C_NRUNS;
    f.cap = cap_runs;
    f.abort = cd.c + C_ABORT;
    scan_launch(f, nullptr, 0, L, pool, s, cd.c + C_ABORT);
}

struct DpOut {
    uint32_t *best_last = nullptr;   // entry index chosen at p = L - 1 when L - 1 is inside a run
    unsigned long long *score_total = nullptr;
};
struct Entry {
    uint16_t bases, delta;
    uint32_t count;
    uint32_t g;  // sparse index or 0xFFFFFFFF for the reference 3-mer
};
__device__ __forceinline__ uint32_t n_ent(const MsaDev &m, uint32_t p) {
    return (p >= 2 ? 1u : 0u) + m.sp_cnt[p];
}
__device__ __forceinline__ Entry get_entry(const MsaDev &m, uint32_t p, uint32_t idx) {
    Entry e;
    if (p >= 2 && idx == 0) {
        e.bases = (uint16_t)((uint32_t)m.code[p - 2] << 8 | (uint32_t)m.code[p - 1] << 4 | m.code[p]);
        e.delta = 0;
        e.count = m.dense_cnt[p];
        e.g = 0xFFFFFFFFu;
    } else {
        e.g = m.sp_off[p] + idx - (p >= 2 ? 1 : 0);
        e.bases = m.g_bases[e.g];
        e.delta = m.g_delta[e.g];
        e.count = m.g_count[e.g];
    }
    return e;
}
constexpr int64_t kDead = INT64_MIN >> 1;  // main.rs:1661

// Score of entry idx of position p inside the run that starts at s (main.rs:1653-1679): best predecessor among the
// entries of b2's position whose (b2, b3) equal this entry's (b1, b2).
__device__ __forceinline__ int64_t dp_entry(const MsaDev &m, uint32_t s, uint32_t p, uint32_t idx, int64_t cov, bool single,
                                            uint32_t &kd_out) {
    const Entry x = get_entry(m, p, idx);
    ABase b1, b2, b3;
    kmer_bases(x.bases, x.delta, p, b1, b2, b3);
    kd_out = kmer_b3delta(x.bases, x.delta);
    const int64_t inc = 10 * (int64_t)x.count - 4 * cov;
    uint32_t besti = 0;
    int64_t score;
    if (b2.q == 15) {
        score = inc;
    } else {
        score = kDead;
        const uint32_t pp = b2.t_pos;
        const uint32_t base23 = ((uint32_t)b1.q << 4 | b2.q) & 255u;
        const uint32_t delta23 = b1.t_pos == b2.t_pos ? 1u : 0u;
        const bool boundary = pp < s;  // the articulation before the run: one entry, local score 0
        const uint32_t npe = boundary ? 1u : (pp == p ? idx : n_ent(m, pp));
        for (uint32_t pi = 0; pi < npe; pi++) {
            const Entry v = get_entry(m, pp, pi);
            if ((v.bases & 255u) != base23 || ((v.bases >> 12) & 1u) != delta23) continue;
            ABase v1, v2, v3;
            kmer_bases(v.bases, v.delta, pp, v1, v2, v3);
            if (!(v2.eq(b1) && v3.eq(b2))) continue;
            if (pp >= 3 && v1.q == 15) continue;  // main.rs:1666-1668
            const int64_t vs = boundary ? 0 : (v.g == 0xFFFFFFFFu ? m.dense_score[pp] : m.g_score[v.g]);
            const int64_t sc = vs + inc;
            if (sc > score || (sc == score && v1.q != 4)) {  // main.rs:1670
                score = sc;
                besti = pi;
            }
        }
    }
    if (x.g == 0xFFFFFFFFu) {
        m.dense_besti[p] = besti;
        if (!single) m.dense_score[p] = score;
    } else {
        m.g_besti[x.g] = besti;
        m.g_score[x.g] = score;
    }
    return score;
}
constexpr uint32_t kDpLongWork = 4096;  // entry x predecessor tests after which a run is handed to a whole warp
                                        // (NP2_DP_LONG_WORK overrides it: 0 sends every run there, for the tests)

// One thread per run of multi-entry positions [s, e); e is the next articulation position (single entry,
// every path passes through it), so scores can be kept relative to the articulation before s.
// A run whose work (entries x predecessors) passes kDpLongWork — a tandem-repeat block with tens of 3-mers per position
// — is put on a queue and done again from its start by a whole warp (k_dp_runs_long).
__global__ void k_dp_runs(MsaDev m, const uint32_t *__restrict__ run_start, DpOut o, uint32_t *__restrict__ long_runs,
                          uint32_t *__restrict__ n_long, uint32_t long_work) {
    uint32_t ri = blockIdx.x * blockDim.x + threadIdx.x;
    if (m.cnt[C_ABORT] || ri >= m.cnt[C_NRUNS]) return;
    const uint32_t s = run_start[ri], L = m.L;
    int64_t best = 0;
    uint32_t best_idx = 0, work = 0, ne_prev = 1;
    for (uint32_t p = s; p < L; p++) {
        const bool single = !m.multi[p];
        const uint32_t ne = n_ent(m, p);
        const int64_t cov = m.cover[p];
        work += ne * ne_prev;
        if (work > long_work) {
            long_runs[atomicAdd(n_long, 1u)] = s;
            return;
        }
        ne_prev = ne;
        for (uint32_t idx = 0; idx < ne; idx++) {
            uint32_t kd;
            const int64_t score = dp_entry(m, s, p, idx, cov, single, kd);
            if (p == L - 1 && (idx == 0 || score >= best)) {  // main.rs:1680, offset-invariant form
                best = score;
                best_idx = idx;
            }
        }
        if (single) return;
    }
    *o.best_last = best_idx;  // the run reached the contig end
}
// One warp per long run.  Positions stay sequential (every position needs the scores of the one before); the entries of a
// position are spread over the lanes.  An entry whose predecessor sits at the SAME position (b3 is an insertion column)
// needs that predecessor's score first: its b3.delta is one less, and the entries of a position are ordered by b3.delta
// (Msa::sort), so the lanes take them in waves of equal b3.delta.
__global__ void __launch_bounds__(128) k_dp_runs_long(MsaDev m, DpOut o, const uint32_t *__restrict__ long_runs,
                                                      const uint32_t *__restrict__ n_long) {
    const uint32_t lane = threadIdx.x & 31, L = m.L;
    if (m.cnt[C_ABORT]) return;
    const uint32_t nl = *n_long, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < nl; w += nw) {
        const uint32_t s = long_runs[w];
        int64_t best = 0;
        uint32_t best_idx = 0;
        bool any = false, ended = false;
        for (uint32_t p = s; p < L && !ended; p++) {
            const bool single = !m.multi[p];
            const uint32_t ne = n_ent(m, p);
            const int64_t cov = m.cover[p];
            for (uint32_t done = 0; done < ne;) {
                const uint32_t idx = done + lane;
                uint32_t kd = 0xFFFFFFFFu;
                if (idx < ne) {
                    const Entry x = get_entry(m, p, idx);
                    kd = kmer_b3delta(x.bases, x.delta);
                }
                const uint32_t kd0 = __shfl_sync(0xFFFFFFFFu, kd, 0);
                const uint32_t same = __ballot_sync(0xFFFFFFFFu, idx < ne && kd == kd0);
                const uint32_t cnt = same == 0xFFFFFFFFu ? 32u : (uint32_t)__ffs(~same) - 1;  // leading lanes of the wave
                int64_t score = INT64_MIN;
                if (lane < cnt) {
                    uint32_t kd_;
                    score = dp_entry(m, s, p, idx, cov, single, kd_);
                }
                __syncwarp();
                if (p == L - 1) {  // main.rs:1680: the last entry among those with the highest score
                    int64_t mx = score;
                    for (int d = 16; d > 0; d >>= 1) mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, d));
                    const uint32_t at = __ballot_sync(0xFFFFFFFFu, lane < cnt && score == mx);
                    if (!any || mx >= best) {
                        best = mx;
                        best_idx = done + (31 - __clz(at));
                        any = true;
                    }
                }
                done += cnt;
            }
            if (single) ended = true;
        }
        if (!ended && lane == 0) *o.best_last = best_idx;  // the run reached the contig end
    }
}
static DpOut dp_out(CountsDev cd) {
    DpOut o;
    o.best_last = cd.c + C_BESTLAST;
    o.score_total = cd.q + Q_TOTAL;
    return o;
}
void dp_runs(MsaDev m, const uint32_t *d_run_start, uint32_t cap_runs, CountsDev cd, uint32_t *d_long_runs, cudaStream_t s) {
    if (!cap_runs) return;
    static const uint32_t long_work = getenv("NP2_DP_LONG_WORK") ? (uint32_t)atoi(getenv("NP2_DP_LONG_WORK")) : kDpLongWork;
    NP2_K(k_dp_runs)<<<cdiv(cap_runs, 64), 64, 0, s>>>(m, d_run_start, dp_out(cd), d_long_runs, cd.c + C_NLONGRUN, long_work);
    NP2_K(k_dp_runs_long)<<<148 * 4, 128, 0, s>>>(m, dp_out(cd), d_long_runs, cd.c + C_NLONGRUN);
}

// Backtrack of one run (main.rs:1572-1634 without the LQ state machine).  WRITE = false: count emitted
// bases per position; WRITE = true: write them at the scanned offsets (ascending order == reversed vec).
template <bool WRITE>
__global__ void k_emit_runs(MsaDev m, const uint32_t *__restrict__ run_start, DpOut o,
                            uint32_t *__restrict__ n_emit, const uint32_t *__restrict__ emit_off,
                            uint32_t *__restrict__ out_pos, uint8_t *__restrict__ out_base,
                            uint8_t *__restrict__ out_flags) {
    uint32_t ri = blockIdx.x * blockDim.x + threadIdx.x;
    long long sum = 0;
    if (m.cnt[C_ABORT]) return;
    const uint32_t n_runs = m.cnt[C_NRUNS];
    if (ri < n_runs) {
        const uint32_t s = run_start[ri], L = m.L;
        uint32_t e = s;
        while (e < L && m.multi[e]) e++;
        uint32_t p, idx;
        if (e < L) {
            p = e - 1;  // b2 of the reference 3-mer at e sits at e - 1
            idx = m.dense_besti[e];
        } else {
            p = L - 1;
            idx = *o.best_last;
        }
        uint32_t cur_p = 0xFFFFFFFFu, cur_n = 0, cur_w = 0;
        for (;;) {
            if (s > 0 && p < s) break;  // reached the articulation before the run
            const Entry x = get_entry(m, p, idx);
            ABase b1, b2, b3;
            kmer_bases(x.bases, x.delta, p, b1, b2, b3);
            const int64_t cov = m.cover[p];
            if (!WRITE) sum += 10 * (long long)x.count - 4 * cov;
            if (p != cur_p) {
                if (!WRITE && cur_p != 0xFFFFFFFFu) n_emit[cur_p] = cur_n;
                cur_p = p;
                cur_n = 0;
                if (WRITE) cur_w = emit_off[p] + n_emit[p];
            }
            if (b3.q != 4) {
                cur_n++;
                if (WRITE) {
                    cur_w--;
                    out_pos[cur_w] = p;
                    out_base[cur_w] = code_char(b3.q);
                    const int64_t qv = (int64_t)x.count * 100 / cov;  // main.rs:1576
                    out_flags[cur_w] = (uint8_t)((qv < 95 ? 1 : 0) | (cov < 2 ? 2 : 0));
                }
            }
            if (b2.q == 15) break;  // main.rs:1629
            idx = x.g == 0xFFFFFFFFu ? m.dense_besti[p] : m.g_besti[x.g];
            p = b2.t_pos;
        }
        if (!WRITE && cur_p != 0xFFFFFFFFu) n_emit[cur_p] = cur_n;
    }
    if (!WRITE) {
        for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, d);
        if ((threadIdx.x & 31) == 0 && sum != 0) atomicAdd(o.score_total, (unsigned long long)sum);
    }
}
// emit_off = exclusive sum of n_emit; a single-entry position emits its own consensus base right here (its only entry is
// the reference 3-mer with count == coverage: qv = 100, so the only flag it can carry is coverage < 2, main.rs:1576-1588)
struct ScanEmit : ScanSumBase {
    MsaDev m;
    const uint32_t *n_emit;
    uint32_t *emit_off, *out_pos;
    uint8_t *out_base, *out_flags;
    uint32_t cap_n;
    uint32_t *count, *abort;
    __device__ uint32_t load(uint32_t p) const { return n_emit[p]; }
    __device__ void store(uint32_t p, unsigned long long ex, unsigned long long in) const {
        emit_off[p] = (uint32_t)ex;
        if (in != ex && ex < cap_n && !m.multi[p]) {
            out_pos[ex] = p;
            out_base[ex] = code_char(m.code[p]);
            out_flags[ex] = m.cover[p] < 2 ? 2 : 0;
        }
    }
    __device__ void total(unsigned long long t, uint32_t n) const {
        emit_off[n] = (uint32_t)t;
        *count = (uint32_t)(t > 0xFFFFFFFFULL ? 0xFFFFFFFFULL : t);
        if (t > cap_n) atomicExch(abort, 1u);
    }
};
void emit_count_runs(MsaDev m, const uint32_t *d_run_start, uint32_t cap_runs, CountsDev cd, uint32_t *d_n_emit,
                     cudaStream_t s) {
    if (!cap_runs) return;
    NP2_K(k_emit_runs<false>)<<<cdiv(cap_runs, 64), 64, 0, s>>>(m, d_run_start, dp_out(cd), d_n_emit, nullptr, nullptr, nullptr,
                                                        nullptr);
}
void emit_offsets(MsaDev m, const uint32_t *d_n_emit, uint32_t *d_emit_off, uint32_t *d_pos, uint8_t *d_base, uint8_t *d_flags,
                  uint32_t cap_n, CountsDev cd, ScanPool &pool, cudaStream_t s) {
    ScanEmit f;
    f.m = m;
    f.n_emit = d_n_emit;
    f.emit_off = d_emit_off;
    f.out_pos = d_pos;
    f.out_base = d_base;
    f.out_flags = d_flags;
    f.cap_n = cap_n;
    f.count = cd.c + C_N;
    f.abort = cd.c + C_ABORT;
    scan_launch(f, nullptr, 0, m.L, pool, s, cd.c + C_ABORT);
}
void emit_write(MsaDev m, const uint32_t *d_run_start, uint32_t cap_runs, CountsDev cd, const uint32_t *d_n_emit,
                const uint32_t *d_emit_off, uint32_t *d_pos, uint8_t *d_base, uint8_t *d_flags, cudaStream_t s) {
    if (cap_runs)
        NP2_K(k_emit_runs<true>)<<<cdiv(cap_runs, 64), 64, 0, s>>>(m, d_run_start, dp_out(cd), const_cast<uint32_t *>(d_n_emit),
                                                           d_emit_off, d_pos, d_base, d_flags);
}
void events_select(const uint8_t *d_cflags, uint32_t cap_n, uint32_t *d_events, uint32_t cap_ev, CountsDev cd,
                   ScanPool &pool, cudaStream_t s) {
    ScanSelect<PredFlagU8> f;
    f.pred = PredFlagU8{d_cflags};
    f.out = d_events;
    f.count = cd.c + C_NEV;
    f.cap = cap_ev;
    f.abort = cd.c + C_ABORT;
    scan_launch(f, cd.c + C_N, 0, cap_n, pool, s, cd.c + C_ABORT);
}

}  // namespace np2
