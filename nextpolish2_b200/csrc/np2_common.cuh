// np2_common.cuh — shared device/host helpers for libnp2gpu (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "np2_error.h"

namespace np2 {

// launch accounting for bench.py's "gpu_launches": every launch of one of OUR kernels goes through NP2_K(...)
unsigned long long &launch_counter();
template <class F>
inline F *count_launch(F *f) {
    ++launch_counter();
    return f;
}
#define NP2_K(k) (*np2::count_launch(k))

#define NP2_CUDA(call)                                                                                  \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            throw np2::Error(-1, std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + \
                                     std::to_string(__LINE__) + ")");                                   \
    } while (0)

// SEQ_NUM (reference src/utils/kmer.rs:11-22) for a byte < 128: A0 C1 G2 T/U3 N5 M6, everything else 4.
__host__ __device__ __forceinline__ uint32_t seq_code(uint32_t c) {
    uint32_t u = c & 0xDFu;  // fold case; only letters can collide with letters
    uint32_t r = 4;
    r = (u == 'A') ? 0u : r;
    r = (u == 'C') ? 1u : r;
    r = (u == 'G') ? 2u : r;
    r = (u == 'T' || u == 'U') ? 3u : r;
    r = (u == 'N') ? 5u : r;
    r = (u == 'M') ? 6u : r;
    return r;
}
// code -> char (first 7 entries of SEQ_NUM)
__host__ __device__ __forceinline__ uint8_t code_char(uint32_t code) {
    const uint64_t tab = 0x004D4E2D54474341ULL;  // "ACGT-NM"
    return (uint8_t)(tab >> (8 * code));
}
// BAM 4-bit base "=ACMGRSVTWYHKDBN" -> SEQ_NUM code of the decoded ASCII char (rust-htslib decodes to
// upper-case ASCII; main.rs:292 then maps through SEQ_NUM)
__host__ __device__ __forceinline__ uint32_t bam4_code(uint32_t b4) {
    // idx: 0'=' 1'A' 2'C' 3'M' 4'G' 5'R' 6'S' 7'V' 8'T' 9'W' 10'Y' 11'H' 12'K' 13'D' 14'B' 15'N'
    const uint64_t t = (4ULL << 0) | (0ULL << 4) | (1ULL << 8) | (6ULL << 12) | (2ULL << 16) | (4ULL << 20) |
                       (4ULL << 24) | (4ULL << 28) | (3ULL << 32) | (4ULL << 36) | (4ULL << 40) | (4ULL << 44) |
                       (4ULL << 48) | (4ULL << 52) | (4ULL << 56) | (5ULL << 60);
    return (uint32_t)(t >> (4 * b4)) & 15u;
}
__host__ __device__ __forceinline__ uint8_t bam4_char(uint32_t b4) {
    // "=ACMGRSVTWYHKDBN" as two 8-byte words (no memory table on the device)
    const uint64_t lo = 0x565352474D43413DULL;  // '=' 'A' 'C' 'M' 'G' 'R' 'S' 'V'
    const uint64_t hi = 0x4E42444B48595754ULL;  // 'T' 'W' 'Y' 'H' 'K' 'D' 'B' 'N'
    return (uint8_t)((b4 < 8 ? lo : hi) >> (8 * (b4 & 7)));
}

// yak hashes (reference src/utils/kmer.rs:223-249, yak/yak-priv.h:10-38)
__host__ __device__ __forceinline__ uint64_t yak_hash64(uint64_t key, uint64_t mask) {
    key = (~key + (key << 21)) & mask;
    key = key ^ key >> 24;
    key = ((key + (key << 3)) + (key << 8)) & mask;
    key = key ^ key >> 14;
    key = ((key + (key << 2)) + (key << 4)) & mask;
    key = key ^ key >> 28;
    key = (key + (key << 31)) & mask;
    return key;
}
__host__ __device__ __forceinline__ uint64_t yak_hash64_64(uint64_t key) {
    key = ~key + (key << 21);
    key = key ^ key >> 24;
    key = (key + (key << 3)) + (key << 8);
    key = key ^ key >> 14;
    key = (key + (key << 2)) + (key << 4);
    key = key ^ key >> 28;
    key = key + (key << 31);
    return key;
}

// One bucket of the yak table (4 x u64 = one 32-byte sector) in ONE 256-bit load (LDG.E.256, sm_100+), read-only path,
// evict-first in L2: a probe never comes back to its bucket, so the line should not push the streaming data of the
// surrounding kernels out of L2.  L2::64B: a miss then fills 64 bytes instead of the whole 128-byte line (measured,
// profiles/microbench/l2_fetch.cu: 61 instead of 117 DRAM bytes per random 32-byte load; the load RATE is the same
// 43 G/s either way — random accesses are bound by DRAM activations, not bytes — so this only takes the needless
// half of the traffic off the memory system; cudaLimitMaxL2FetchGranularity does not change either number).
__device__ __forceinline__ void ld_bucket(const uint64_t *p, uint64_t v[4]) {
    asm volatile("ld.global.nc.L2::evict_first.L2::64B.v4.u64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3])
                 : "l"(p));
}

// One aligned base (reference AlignBase main.rs:33-52), packed for registers.
struct ABase {
    uint32_t t_pos;
    uint16_t delta;
    uint8_t q;
    __host__ __device__ bool eq(const ABase &o) const { return t_pos == o.t_pos && delta == o.delta && q == o.q; }
};
// Kmer::bases(p) (main.rs:105-184); u32/u16 arithmetic wraps like the release build.
__host__ __device__ __forceinline__ void kmer_bases(uint16_t bases, uint16_t delta, uint32_t p, ABase &a, ABase &b,
                                                    ABase &c) {
    a.q = (bases >> 8) & 15;
    b.q = (bases >> 4) & 15;
    c.q = bases & 15;
    a.delta = delta;
    if ((bases & 0x5000) == 0x5000) {  // A--
        a.t_pos = p;
        b.t_pos = p;
        b.delta = (uint16_t)(delta + 1);
        c.t_pos = p;
        c.delta = (uint16_t)(delta + 2);
    } else if (bases & 0x1000) {  // AA-
        a.t_pos = p - 1;
        b.t_pos = p;
        b.delta = 0;
        c.t_pos = p;
        c.delta = 1;
    } else if (bases & 0x4000) {  // A-A
        a.t_pos = p - 1;
        b.t_pos = p - 1;
        b.delta = (uint16_t)(delta + 1);
        c.t_pos = p;
        c.delta = 0;
    } else {  // AAA
        a.t_pos = p - 2;
        b.t_pos = p - 1;
        b.delta = 0;
        c.t_pos = p;
        c.delta = 0;
    }
}
__host__ __device__ __forceinline__ uint16_t kmer_b3delta(uint16_t bases, uint16_t delta) {
    if ((bases & 0x5000) == 0x5000) return (uint16_t)(delta + 2);
    if (bases & 0x1000) return 1;
    return 0;
}

}  // namespace np2
