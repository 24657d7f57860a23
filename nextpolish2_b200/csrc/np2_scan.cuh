// np2_scan.cuh — single-pass prefix scans and stream compaction for the polish pipeline (sm_100a).
//
// Every "how many / where does it go" step of the pipeline (records -> groups, positions -> runs, consensus offsets,
// LQ events -> regions, (read, region) pairs, pool offsets ...) is a prefix scan whose element count is itself the
// result of an earlier kernel.  These scans therefore read their element count from DEVICE memory (the grid is sized
// from a host-side capacity) and hand their total back to device memory, so that the host never has to synchronise
// just to size the next launch.  The element producer is fused in: a scan is parameterised by a small functor
// (load / op / store / total), so "flag, then scan the flags, then scatter" is one kernel.
//
// Algorithm: one pass, chained tiles with decoupled look-back.  A tile is kScanThreads x kScanItems elements; tiles
// take their index from an atomic ticket (so a tile only ever waits for tiles that already run); every tile publishes
// one 64-bit descriptor {status:2, value:62} (aggregate first, inclusive prefix later) and a warp walks 32
// predecessors at a time.  Values are kept in 62 bits: sums are taken modulo 2^62 (functors sign-extend if they scan
// signed values), which every count / offset of this pipeline fits by a wide margin.
#pragma once
#include "np2_common.cuh"

namespace np2 {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;
constexpr unsigned long long kScanMask = (1ULL << 62) - 1;

// zeroed 64-bit words for the tickets and tile descriptors of the scans of one pipeline pass; slices are handed out in
// enqueue order and never reused inside a pass, the whole used prefix is cleared by ONE memset at the start of the next
struct ScanPool {
    unsigned long long *d = nullptr;
    size_t cap = 0, used = 0, dirty = 0;
    void reserve(size_t words, cudaStream_t s) {
        if (words <= cap) return;
        if (d) cudaFreeAsync(d, s);
        cap = words + words / 2;
        NP2_CUDA(cudaMallocAsync((void **)&d, cap * 8, s));
        NP2_CUDA(cudaMemsetAsync(d, 0, cap * 8, s));
        used = dirty = 0;
    }
    void begin(cudaStream_t s) {  // start of a pass: everything handed out before is dead
        dirty = used > dirty ? used : dirty;
        if (dirty) NP2_CUDA(cudaMemsetAsync(d, 0, dirty * 8, s));
        used = dirty = 0;
    }
    unsigned long long *take(size_t words, cudaStream_t s) {
        if (used + words > cap) {  // grow: the old block stays valid for the kernels already enqueued (stream order)
            unsigned long long *old = d;
            d = nullptr;
            const size_t want = (used + words) * 2 + 4096;
            cap = 0;
            reserve(want, s);
            if (old) cudaFreeAsync(old, s);
        }
        unsigned long long *p = d + used;
        used += words;
        return p;
    }
    void destroy(cudaStream_t s) {
        if (d) cudaFreeAsync(d, s);
        d = nullptr;
        cap = used = dirty = 0;
    }
};

__device__ __forceinline__ unsigned long long scan_ld(const unsigned long long *p) {
    return *reinterpret_cast<const volatile unsigned long long *>(p);
}
__device__ __forceinline__ void scan_st(unsigned long long *p, unsigned long long v) {
    *reinterpret_cast<volatile unsigned long long *>(p) = v;
}

// Tr: struct with
//   __device__ unsigned long long load(uint32_t i) const        value of element i (< n), already reduced to 62 bits
//   __device__ static unsigned long long op(a, b)               associative and commutative, closed on 62-bit values
//   __device__ static unsigned long long identity()
//   __device__ void store(uint32_t i, excl, incl) const         prefix before / including element i
//   __device__ void total(unsigned long long t, uint32_t n) const   once, after the last element (also when n == 0)
// n = *d_n + n_plus (clamped to cap) when d_n != nullptr, else cap.
template <class Tr>
__global__ void __launch_bounds__(kScanThreads) k_scan(Tr tr, const uint32_t *__restrict__ d_n, uint32_t n_plus, uint32_t cap,
                                                       unsigned long long *ws, const uint32_t *__restrict__ d_abort) {
    __shared__ unsigned long long s_warp[kScanThreads / 32], s_prefix;
    __shared__ uint32_t s_tile;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (d_abort && *d_abort) return;  // set by an earlier kernel: the same for every tile of this launch
    if (tid == 0) s_tile = atomicAdd(reinterpret_cast<unsigned int *>(ws), 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    uint32_t n = cap;
    if (d_n) {
        const unsigned long long want = (unsigned long long)*d_n + n_plus;
        n = want < cap ? (uint32_t)want : cap;
    }
    const uint64_t base64 = (uint64_t)tile * kScanTile;
    if (base64 >= n && tile > 0) return;
    unsigned long long *desc = ws + 1;
    // Warp-striped: a warp owns kScanItems * 32 consecutive elements, element j * 32 + lane of them sits in loc[j] of
    // lane `lane`, so every load and store instruction of a warp touches 32 consecutive elements.
    const uint64_t w0 = base64 + (uint64_t)warp * (kScanItems * 32) + lane;
    unsigned long long loc[kScanItems];
#pragma unroll
    for (int j = 0; j < kScanItems; j++) {
        const uint64_t i = w0 + (uint64_t)j * 32;
        loc[j] = i < n ? tr.load((uint32_t)i) : Tr::identity();
    }
    // inclusive scan in element order inside the warp: a shuffle scan per row, rows chained through lane 31
    unsigned long long carry = Tr::identity();
#pragma unroll
    for (int j = 0; j < kScanItems; j++) {
        unsigned long long v = loc[j];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, v, d);
            if (lane >= d) v = Tr::op(t, v);
        }
        v = Tr::op(carry, v);
        loc[j] = v;
        carry = __shfl_sync(0xFFFFFFFFu, v, 31);
    }
    if (lane == 31) s_warp[warp] = carry;
    __syncthreads();
    unsigned long long warp_ex = Tr::identity(), agg = Tr::identity();
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; w++) {
        if (w < (int)warp) warp_ex = Tr::op(warp_ex, s_warp[w]);
        agg = Tr::op(agg, s_warp[w]);
    }
    if (warp == 0) {
        unsigned long long ex = Tr::identity();
        if (tile == 0) {
            if (lane == 0) scan_st(desc, 2ULL << 62 | (agg & kScanMask));
        } else {
            if (lane == 0) scan_st(desc + tile, 1ULL << 62 | (agg & kScanMask));
            int look = (int)tile - 1;
            for (;;) {
                const int idx = look - (int)lane;
                unsigned long long dsc = idx >= 0 ? scan_ld(desc + idx) : (2ULL << 62 | (Tr::identity() & kScanMask));
                // wait for the descriptors in front of the nearest inclusive prefix (all of them if there is none yet)
                for (;;) {
                    const uint32_t inc = __ballot_sync(0xFFFFFFFFu, (dsc >> 62) == 2);
                    const uint32_t inv = __ballot_sync(0xFFFFFFFFu, (dsc >> 62) == 0);
                    const uint32_t need = inc ? ((inc & (0u - inc)) - 1u) : 0xFFFFFFFFu;  // lanes below the first inclusive one
                    if (!(inv & need)) break;
                    __nanosleep(40);
                    if ((dsc >> 62) == 0) dsc = scan_ld(desc + idx);
                }
                const uint32_t m = __ballot_sync(0xFFFFFFFFu, (dsc >> 62) == 2);
                const uint32_t first = m ? (uint32_t)__ffs(m) - 1 : 32u;
                unsigned long long v = lane <= first ? (dsc & kScanMask) : Tr::identity();
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) v = Tr::op(v, __shfl_xor_sync(0xFFFFFFFFu, v, d));
                ex = Tr::op(v, ex);
                if (m) break;
                look -= 32;
            }
            if (lane == 0) scan_st(desc + tile, 2ULL << 62 | (Tr::op(ex, agg) & kScanMask));
        }
        if (lane == 0) s_prefix = ex;
    }
    __syncthreads();
    const unsigned long long pre = Tr::op(s_prefix, warp_ex) & kScanMask;
    unsigned long long row_carry = Tr::identity();  // inclusive value of the last element of the previous row
#pragma unroll
    for (int j = 0; j < kScanItems; j++) {
        unsigned long long prev = __shfl_up_sync(0xFFFFFFFFu, loc[j], 1);
        if (lane == 0) prev = row_carry;
        row_carry = __shfl_sync(0xFFFFFFFFu, loc[j], 31);
        const uint64_t i = w0 + (uint64_t)j * 32;
        if (i < n) tr.store((uint32_t)i, Tr::op(pre, prev) & kScanMask, Tr::op(pre, loc[j]) & kScanMask);
    }
    const uint32_t last_tile = n ? (n - 1) / kScanTile : 0;
    if (tile == last_tile && tid == 0) tr.total(Tr::op(s_prefix, agg) & kScanMask, n);
}

inline uint32_t scan_tiles(uint32_t cap) { return cap ? (cap + kScanTile - 1) / kScanTile : 1u; }

template <class Tr>
inline void scan_launch(const Tr &tr, const uint32_t *d_n, uint32_t n_plus, uint32_t cap, ScanPool &pool, cudaStream_t s,
                        const uint32_t *d_abort = nullptr) {
    const uint32_t tiles = scan_tiles(cap);
    unsigned long long *ws = pool.take((size_t)tiles + 1, s);
    NP2_K(k_scan<Tr>)<<<tiles, kScanThreads, 0, s>>>(tr, d_n, n_plus, cap, ws, d_abort);
}

/* ---------------------------------------------------------------- the functors the pipeline uses */

struct ScanSumBase {
    __device__ static unsigned long long op(unsigned long long a, unsigned long long b) { return (a + b) & kScanMask; }
    __device__ static unsigned long long identity() { return 0; }
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
    __device__ unsigned long long load(uint32_t i) const { return (unsigned long long)in[i]; }
    __device__ void store(uint32_t i, unsigned long long ex, unsigned long long) const { out[i] = (Out)ex; }
    __device__ void total(unsigned long long t, uint32_t n) const {
        out[n] = (Out)t;
        if (c_slot) *c_slot = (uint32_t)(t > 0xFFFFFFFFULL ? 0xFFFFFFFFULL : t);
        if (q_slot) *q_slot = t;
        if (abort && t > cap) atomicExch(abort, 1u);
    }
};
// signed 64-bit version (patch length deltas)
struct ScanOffsetsI64 : ScanSumBase {
    const long long *in;
    long long *out;
    unsigned long long *q_slot;
    __device__ unsigned long long load(uint32_t i) const { return (unsigned long long)in[i] & kScanMask; }
    __device__ void store(uint32_t i, unsigned long long ex, unsigned long long) const { out[i] = scan_signed(ex); }
    __device__ void total(unsigned long long t, uint32_t n) const {
        out[n] = scan_signed(t);
        if (q_slot) *q_slot = (unsigned long long)scan_signed(t);
    }
};
// in-place inclusive sum of signed 32-bit values (coverage from the difference array)
struct ScanInclusiveI32 : ScanSumBase {
    int32_t *a;
    __device__ unsigned long long load(uint32_t i) const { return (unsigned long long)(long long)a[i] & kScanMask; }
    __device__ void store(uint32_t i, unsigned long long, unsigned long long in) const { a[i] = (int32_t)scan_signed(in); }
    __device__ void total(unsigned long long, uint32_t) const {}
};
// in-place inclusive minimum (the region cursor only moves down, main.rs:1446-1448)
struct ScanInclusiveMinU32 {
    uint32_t *a;
    __device__ static unsigned long long op(unsigned long long x, unsigned long long y) { return x < y ? x : y; }
    __device__ static unsigned long long identity() { return kScanMask; }
    __device__ unsigned long long load(uint32_t i) const { return a[i]; }
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
    __device__ unsigned long long load(uint32_t i) const { return pred(i) ? 1ULL : 0ULL; }
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
