// np2_scan.cuh — prefix scans and stream compaction for the polish pipeline (sm_100a).
//
// Every "how many / where does it go" step of the pipeline (positions -> runs, consensus offsets, LQ events ->
// regions, (read, region) pairs, pool offsets ...) is a prefix scan whose element count is itself the result of an
// earlier kernel.  These scans therefore read their element count from DEVICE memory (the grid is sized from a
// host-side capacity) and hand their total back to device memory, so that the host never has to synchronise just to
// size the next launch.  The element producer / consumer is fused in: a scan is parameterised by a small functor
// (load / op / store / total), so "flag, scan the flags, scatter" or "count, scan, emit" is one scan.
//
// Algorithm: reduce-then-scan over a FIXED partition.  The elements are cut into one contiguous chunk per CTA
// (at most kScanMaxCtas of them); pass 1 reduces every chunk (loads only) and its last CTA to finish scans the chunk
// totals; pass 2 re-reads every chunk, scans it in registers (every thread owns 16 consecutive elements: 15 serial operations, one
// shuffle scan per warp) and calls the functor's store.  The input is read twice, but nothing ever waits for
// another CTA: a chained single-pass scan with look-back measured 2-3x slower here, because with ~500 tiles in flight
// every tile walks back through hundreds of descriptors that only hold aggregates (profiles/r02h_launches.csv).
// Small inputs (one chunk) take a single launch.  Values are kept in 62 bits (functors sign-extend if they scan signed
// values), which every count / offset of this pipeline fits by a wide margin.
#pragma once
#include "np2_common.cuh"

namespace np2 {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;
constexpr int kScanMaxCtas = 592;  // 4 per SM
constexpr int kScanSingleCtaTiles = 8;  // up to 32768 elements are scanned by one CTA in one launch (three launches cost more)
constexpr unsigned long long kScanMask = (1ULL << 62) - 1;

// 64-bit scratch words for the chunk totals of the scans of one pipeline pass; slices are handed out in enqueue order
// and never reused inside a pass
constexpr int kScanTickets = 64;
struct ScanPool {
    unsigned long long *d = nullptr;
    size_t cap = 0, used = 0;
    // "which CTA of the reduce pass finishes last" counters: zero when a scan starts, set back to zero by the CTA that
    // draws the last ticket.  The scans of a pool run on one stream, one after the other; they still rotate through a
    // few counters so that no scan depends on the one before it having cleaned up.
    uint32_t *tickets = nullptr;
    uint32_t n_scans = 0;
    void reserve(size_t words, cudaStream_t s) {
        if (!tickets) {
            NP2_CUDA(cudaMallocAsync((void **)&tickets, kScanTickets * 4, s));
            NP2_CUDA(cudaMemsetAsync(tickets, 0, kScanTickets * 4, s));
        }
        if (words <= cap) return;
        if (d) cudaFreeAsync(d, s);
        cap = words + words / 2;
        NP2_CUDA(cudaMallocAsync((void **)&d, cap * 8, s));
        used = 0;
    }
    uint32_t *ticket(cudaStream_t s) {
        if (!tickets) reserve(1, s);
        return tickets + (n_scans++ % kScanTickets);
    }
    void begin(cudaStream_t) { used = 0; }  // start of a pass: everything handed out before is dead (same stream)
    unsigned long long *take(size_t words, cudaStream_t s) {
        if (used + words > cap) {  // grow: the old block stays valid for the kernels already enqueued (stream order)
            unsigned long long *old = d;
            d = nullptr;
            cap = 0;
            reserve((used + words) * 2 + 4096, s);
            if (old) cudaFreeAsync(old, s);
        }
        unsigned long long *p = d + used;
        used += words;
        return p;
    }
    void destroy(cudaStream_t s) {
        if (d) cudaFreeAsync(d, s);
        if (tickets) cudaFreeAsync(tickets, s);
        d = nullptr;
        tickets = nullptr;
        cap = used = 0;
    }
};

// Tr: struct with
//   __device__ uint32_t load(uint32_t i) const                  value of element i (< n); no side effects (it is called
//                                                               twice per element)
//   __device__ static uint32_t op32(a, b), identity32()         the operation on 32-bit values: used INSIDE a tile of 4096
//                                                               elements, whose partial results must fit 32 bits
//   __device__ static unsigned long long widen(uint32_t)        a tile-local partial result as a 62-bit value
//   __device__ static unsigned long long op(a, b), identity()   the operation on 62-bit values (chunk and tile prefixes);
//                                                               associative and commutative
//   __device__ void store(uint32_t i, excl, incl) const         prefix before / including element i (62-bit values)
//   __device__ void total(unsigned long long t, uint32_t n) const   once, after the last element (also when n == 0)
// n = *d_n + n_plus (clamped to cap) when d_n != nullptr, else cap.
__device__ __forceinline__ uint32_t scan_count(const uint32_t *d_n, uint32_t n_plus, uint32_t cap) {
    if (!d_n) return cap;
    const unsigned long long want = (unsigned long long)*d_n + n_plus;
    return want < cap ? (uint32_t)want : cap;
}
// chunk of CTA c: [c * chunk, min(n, (c + 1) * chunk)), chunk a multiple of the tile
__device__ __forceinline__ uint64_t scan_chunk(uint32_t n, uint32_t ctas) {
    const uint64_t per = ((uint64_t)n + ctas - 1) / ctas;
    return (per + kScanTile - 1) / kScanTile * kScanTile;
}
template <class Tr>
__device__ __forceinline__ unsigned long long scan_block_reduce(unsigned long long v, unsigned long long *s_warp) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = Tr::op(v, __shfl_xor_sync(0xFFFFFFFFu, v, d));
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned long long r = Tr::identity();
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; w++) r = Tr::op(r, s_warp[w]);
    return r;
}

// one CTA's worth of work: exclusive scan of the chunk totals in place, the grand total goes to the functor
template <class Tr>
__device__ __forceinline__ void scan_partials(const Tr &tr, const uint32_t *__restrict__ d_n, uint32_t n_plus, uint32_t cap,
                                              unsigned long long *partial, uint32_t ctas, unsigned long long *s_warp) {
    constexpr int kPer = (kScanMaxCtas + kScanThreads - 1) / kScanThreads;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned long long v[kPer], sum = Tr::identity();
#pragma unroll
    for (int u = 0; u < kPer; u++) {
        const uint32_t c = tid * kPer + u;
        v[u] = c < ctas ? __ldcg(partial + c) : Tr::identity();  // written by other CTAs: not through this SM's L1
        sum = Tr::op(sum, v[u]);
    }
    unsigned long long incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl = Tr::op(t, incl);
    }
    __syncthreads();  // s_warp may still be read by the block reduction before
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned long long ex = Tr::identity(), all = Tr::identity();
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; w++) {
        if (w < (int)warp) ex = Tr::op(ex, s_warp[w]);
        all = Tr::op(all, s_warp[w]);
    }
    unsigned long long lane_ex = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
    if (lane == 0) lane_ex = Tr::identity();
    ex = Tr::op(ex, lane_ex);
#pragma unroll
    for (int u = 0; u < kPer; u++) {
        const uint32_t c = tid * kPer + u;
        if (c < ctas) partial[c] = ex & kScanMask;
        ex = Tr::op(ex, v[u]);
    }
    if (tid == 0) tr.total(all & kScanMask, scan_count(d_n, n_plus, cap));
}
// pass 1: every CTA reduces its chunk; the CTA that finishes LAST (a ticket counter) also scans the chunk totals, so
// the one-CTA pass in between costs no launch of its own
template <class Tr>
__global__ void __launch_bounds__(kScanThreads) k_scan_reduce(Tr tr, const uint32_t *__restrict__ d_n, uint32_t n_plus,
                                                              uint32_t cap, unsigned long long *partial,
                                                              uint32_t *ticket, const uint32_t *__restrict__ d_abort) {
    __shared__ unsigned long long s_warp[kScanThreads / 32];
    __shared__ uint32_t s_last;
    if (d_abort && *d_abort) return;
    const uint32_t n = scan_count(d_n, n_plus, cap);
    const uint64_t chunk = scan_chunk(n, gridDim.x), begin = blockIdx.x * chunk;
    const uint64_t end = begin + chunk < n ? begin + chunk : n;
    unsigned long long acc = Tr::identity();
    for (uint64_t i = begin + threadIdx.x; i < end; i += kScanThreads) acc = Tr::op(acc, Tr::widen(tr.load((uint32_t)i)));
    acc = scan_block_reduce<Tr>(acc, s_warp);
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = acc & kScanMask;
        __threadfence();  // the total is visible before the ticket is drawn
        const uint32_t t = atomicAdd(ticket, 1u);
        s_last = t == gridDim.x - 1;
        if (s_last) *ticket = 0;  // for the next scan that uses this counter (same stream: after this kernel)
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    scan_partials(tr, d_n, n_plus, cap, partial, gridDim.x, s_warp);
}
// pass 2 (or the only pass when there is one chunk: partial == nullptr, the total is reported here)
template <class Tr>
__global__ void __launch_bounds__(kScanThreads) k_scan_apply(Tr tr, const uint32_t *__restrict__ d_n, uint32_t n_plus,
                                                             uint32_t cap, const unsigned long long *__restrict__ partial,
                                                             const uint32_t *__restrict__ d_abort) {
    __shared__ unsigned long long s_warp[kScanThreads / 32];
    if (d_abort && *d_abort) return;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n = scan_count(d_n, n_plus, cap);
    const uint64_t chunk = scan_chunk(n, gridDim.x), begin = blockIdx.x * chunk;
    const uint64_t end = begin + chunk < n ? begin + chunk : n;
    __shared__ uint32_t s_w32[kScanThreads / 32];
    unsigned long long tile_pre = partial ? partial[blockIdx.x] : Tr::identity();
    for (uint64_t t0 = begin; t0 < end; t0 += kScanTile) {
        // Warp-striped: a warp owns kScanItems * 32 consecutive elements, element j * 32 + lane of them sits in loc[j] of
        // lane `lane`, so every load and store instruction of a warp touches 32 consecutive elements.  Inside the tile
        // the arithmetic is 32-bit; the 62-bit prefix of the tile is applied when the results are handed to the functor.
        const uint64_t w0 = t0 + (uint64_t)warp * (kScanItems * 32) + lane;
        uint32_t loc[kScanItems];
#pragma unroll
        for (int j = 0; j < kScanItems; j++) {
            const uint64_t i = w0 + (uint64_t)j * 32;
            loc[j] = i < end ? tr.load((uint32_t)i) : Tr::identity32();
        }
        // inclusive scan of every row (independent shuffle chains), then the rows are chained through their last lane
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
            for (int j = 0; j < kScanItems; j++) {
                const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, loc[j], d);
                if (lane >= d) loc[j] = Tr::op32(t, loc[j]);
            }
        }
        uint32_t carry = Tr::identity32();
#pragma unroll
        for (int j = 0; j < kScanItems; j++) {
            loc[j] = Tr::op32(carry, loc[j]);
            carry = __shfl_sync(0xFFFFFFFFu, loc[j], 31);
        }
        __syncthreads();  // s_w32 of the previous tile has been read
        if (lane == 31) s_w32[warp] = carry;
        __syncthreads();
        uint32_t warp_ex = Tr::identity32(), agg = Tr::identity32();
#pragma unroll
        for (int w = 0; w < kScanThreads / 32; w++) {
            if (w < (int)warp) warp_ex = Tr::op32(warp_ex, s_w32[w]);
            agg = Tr::op32(agg, s_w32[w]);
        }
        uint32_t row_carry = Tr::identity32();  // inclusive value of the last element of the previous row
#pragma unroll
        for (int j = 0; j < kScanItems; j++) {
            uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, loc[j], 1);
            if (lane == 0) prev = row_carry;
            row_carry = __shfl_sync(0xFFFFFFFFu, loc[j], 31);
            const uint64_t i = w0 + (uint64_t)j * 32;
            if (i < end)
                tr.store((uint32_t)i, Tr::op(tile_pre, Tr::widen(Tr::op32(warp_ex, prev))) & kScanMask,
                         Tr::op(tile_pre, Tr::widen(Tr::op32(warp_ex, loc[j]))) & kScanMask);
        }
        tile_pre = Tr::op(tile_pre, Tr::widen(agg)) & kScanMask;
    }
    if (!partial && tid == 0) tr.total(tile_pre, n);
}

inline uint32_t scan_ctas(uint32_t cap) {
    const uint32_t t = cap ? (cap + kScanTile - 1) / kScanTile : 1u;
    return t < (uint32_t)kScanMaxCtas ? t : (uint32_t)kScanMaxCtas;
}

template <class Tr>
inline void scan_launch(const Tr &tr, const uint32_t *d_n, uint32_t n_plus, uint32_t cap, ScanPool &pool, cudaStream_t s,
                        const uint32_t *d_abort = nullptr) {
    uint32_t ctas = scan_ctas(cap);
    if (ctas <= kScanSingleCtaTiles) ctas = 1;  // a few tiles: one CTA walks them in a single launch
    if (ctas == 1) {
        NP2_K(k_scan_apply<Tr>)<<<1, kScanThreads, 0, s>>>(tr, d_n, n_plus, cap, nullptr, d_abort);
        return;
    }
    unsigned long long *partial = pool.take(ctas, s);
    NP2_K(k_scan_reduce<Tr>)<<<ctas, kScanThreads, 0, s>>>(tr, d_n, n_plus, cap, partial, pool.ticket(s), d_abort);
    NP2_K(k_scan_apply<Tr>)<<<ctas, kScanThreads, 0, s>>>(tr, d_n, n_plus, cap, partial, d_abort);
}

/* ---------------------------------------------------------------- the functors the pipeline uses */

struct ScanSumBase {  // sums of unsigned values
    __device__ static unsigned long long op(unsigned long long a, unsigned long long b) { return (a + b) & kScanMask; }
    __device__ static unsigned long long identity() { return 0; }
    __device__ static uint32_t op32(uint32_t a, uint32_t b) { return a + b; }
    __device__ static uint32_t identity32() { return 0; }
    __device__ static unsigned long long widen(uint32_t v) { return v; }
};
struct ScanSignedSumBase : ScanSumBase {  // sums of signed values: tile-local partial sums fit an int32
    __device__ static unsigned long long widen(uint32_t v) { return (unsigned long long)(long long)(int32_t)v & kScanMask; }
};
__device__ __forceinline__ long long scan_signed(unsigned long long v) { return (long long)(v << 2) >> 2; }

// out[i] = sum of in[0..i) for i < n, out[n] = total (the "n + 1 entries" offset array).  The total also goes to
// *c_slot / *q_slot (the device-resident counts) and raises *abort when it exceeds `cap`.
template <class In, class Out>
struct ScanOffsets : ScanSumBase {
    const In *in;
    Out *out;
    uint32_t *c_slot;
    unsigned long long *q_slot;
    unsigned long long cap;
    uint32_t *abort;
    __device__ uint32_t load(uint32_t i) const { return (uint32_t)in[i]; }
    __device__ void store(uint32_t i, unsigned long long ex, unsigned long long) const { out[i] = (Out)ex; }
    __device__ void total(unsigned long long t, uint32_t n) const {
        out[n] = (Out)t;
        if (c_slot) *c_slot = (uint32_t)(t > 0xFFFFFFFFULL ? 0xFFFFFFFFULL : t);
        if (q_slot) *q_slot = t;
        if (abort && t > cap) atomicExch(abort, 1u);
    }
};
// signed 64-bit version (patch length deltas)
struct ScanOffsetsI64 : ScanSignedSumBase {
    const long long *in;
    long long *out;
    unsigned long long *q_slot;
    __device__ uint32_t load(uint32_t i) const { return (uint32_t)in[i]; }
    __device__ void store(uint32_t i, unsigned long long ex, unsigned long long) const { out[i] = scan_signed(ex); }
    __device__ void total(unsigned long long t, uint32_t n) const {
        out[n] = scan_signed(t);
        if (q_slot) *q_slot = (unsigned long long)scan_signed(t);
    }
};
// in-place inclusive sum of signed 32-bit values (coverage from the difference array)
struct ScanInclusiveI32 : ScanSignedSumBase {
    int32_t *a;
    __device__ uint32_t load(uint32_t i) const { return (uint32_t)a[i]; }
    __device__ void store(uint32_t i, unsigned long long, unsigned long long in) const { a[i] = (int32_t)scan_signed(in); }
    __device__ void total(unsigned long long, uint32_t) const {}
};
// in-place inclusive minimum (the region cursor only moves down, main.rs:1446-1448)
struct ScanInclusiveMinU32 {
    uint32_t *a;
    __device__ static unsigned long long op(unsigned long long x, unsigned long long y) { return x < y ? x : y; }
    __device__ static unsigned long long identity() { return kScanMask; }
    __device__ static uint32_t op32(uint32_t x, uint32_t y) { return x < y ? x : y; }
    __device__ static uint32_t identity32() { return 0xFFFFFFFFu; }
    __device__ static unsigned long long widen(uint32_t v) { return v; }
    __device__ uint32_t load(uint32_t i) const { return a[i]; }
    __device__ void store(uint32_t i, unsigned long long, unsigned long long in) const { a[i] = (uint32_t)in; }
    __device__ void total(unsigned long long, uint32_t) const {}
};
// compaction: out[rank] = i for every i with pred(i), *count = how many (may exceed `cap`: then nothing is written past
// the capacity and *abort is raised).  Pred: __device__ bool operator()(uint32_t) const
template <class Pred>
struct ScanSelect : ScanSumBase {
    Pred pred;
    uint32_t *out, *count;
    uint32_t cap;
    uint32_t *abort;
    __device__ uint32_t load(uint32_t i) const { return pred(i) ? 1u : 0u; }
    __device__ void store(uint32_t i, unsigned long long ex, unsigned long long in) const {
        if (in != ex && ex < cap) out[ex] = i;
    }
    __device__ void total(unsigned long long t, uint32_t) const {
        *count = (uint32_t)t;
        if (abort && t > cap) atomicExch(abort, 1u);
    }
};
struct PredFlagU8 {
    const uint8_t *flag;
    __device__ bool operator()(uint32_t i) const { return flag[i] != 0; }
};
struct PredNonZeroU64 {
    const unsigned long long *v;
    __device__ bool operator()(uint32_t i) const { return v[i] != 0; }
};

}  // namespace np2
