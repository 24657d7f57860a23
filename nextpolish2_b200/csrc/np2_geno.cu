// np2_geno.cu — per-region kernels: which reads cover which LQ region, candidate selection (first 60 non-empty in
// read order), candidate k-mer scoring, the genotype rules and the read x read agreement edges.
//
//   k_read_ranges      the monotone region cursor of generate_lqseqs_from_tags_kmer (main.rs:1446-1460)
//   k_region_select    per-region candidate list in read order with the 60 cap (main.rs:1473-1521)
//   k_cand_write       candidate strings into a compact pool
//   k_cand_kscore_*    retrieve_kmer_count (main.rs:740-778)
//   k_region_hete      fill_order_stat + mark_hete_lqseqs (main.rs:813-849, 916-946) + edge counts
//   k_edges_accum      the pair loop of phase_reads_by_lqseqs (main.rs:953-992) into a dense pair accumulator
//   k_phase_*          level 0 of the Louvain graph: ref pairs, `dif <= -3`, invalid reads, CSR (main.rs:972-1010)
//   k_region_seed      fill_order_stat + fill_seed_lqseqs + retain_sort_seqs (main.rs:862-914, 714-726)
#include "np2_kernels.cuh"

namespace np2 {

namespace {
inline uint32_t cdiv(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }
constexpr int kWarpsPerCta = 4;
}  // namespace

/* ---------------------------------------------------------------- read -> region ranges */

// g[i] = the cursor position read i would move an unconstrained cursor to: the largest region index whose start
// is >= t_s (regions are stored in descending position), 0 if there is none.  Blank reads do not move the cursor.
__global__ void k_read_cursor(GenoDev g, ReadsDev R, const uint8_t *__restrict__ blank) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nreg = g.cnt[C_NREG];
    if (i >= R.n_reads || g.cnt[C_ABORT]) return;
    uint32_t v = nreg ? nreg - 1 : 0;
    if (!blank[i] && nreg) {
        const uint32_t ts = R.t_s[i];
        uint32_t lo = 0, hi = nreg;  // count of regions with start >= ts
        while (lo < hi) {
            uint32_t mid = (lo + hi) >> 1;
            if (g.start[mid] >= ts) lo = mid + 1;
            else hi = mid;
        }
        v = lo ? lo - 1 : 0;
    }
    g.rd_s[i] = v;
}
// after the prefix-min over reads: j, pair count and decode limit of every read (main.rs:1449-1471)
__global__ void k_read_ranges(GenoDev g, ReadsDev R, const uint8_t *__restrict__ blank, uint32_t k) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nreg = g.cnt[C_NREG];
    if (i >= R.n_reads || g.cnt[C_ABORT]) return;
    uint32_t np = 0, j = 0;
    if (!blank[i] && nreg) {
        const uint32_t s = g.rd_s[i], ts = R.t_s[i], te = R.t_e[i];
        if (!(g.start[s] < ts || g.end[s] > te)) {
            uint32_t lo = 0, hi = nreg;  // count of regions with end > te
            while (lo < hi) {
                uint32_t mid = (lo + hi) >> 1;
                if (g.end[mid] > te) lo = mid + 1;
                else hi = mid;
            }
            j = lo;
            np = s - j + 1;
        }
    }
    g.rd_j[i] = j;
    g.rd_np[i] = np;
}

/* ---------------------------------------------------------------- candidate scan (one thread per pair) */

struct ScanOut {
    uint32_t len;
    uint64_t kmer;
};
// The read's bases over [start, end] and its first-k canonical k-mer from `start` on (main.rs:1478-1521); the read
// is only decoded up to the first column whose t_pos exceeds `limit` (main.rs:1465-1471).
template <bool WRITE>
__device__ __forceinline__ ScanOut scan_read_region(const ReadsDev &R, uint32_t r, uint32_t start, uint32_t end,
                                                    uint32_t limit, uint32_t k, uint8_t *out) {
    const uint32_t n = R.n[r];
    const uint8_t *nib = R.nib + R.nib_off[r];
    const uint32_t *ck = R.ck_tpos + R.ck_off[r];
    // last 32-column block whose first t_pos is < start (or block 0): guessed from the distance to the read's start
    // (few indels in a HiFi alignment), corrected by a few steps along the checkpoints, else a binary search
    const uint32_t nblk = (n + 31) >> 5, ts = R.t_s[r];
    uint32_t lo = start > ts ? min((start - ts) >> 5, nblk - 1) : 0, steps = 0;
    while (lo > 0 && ck[lo] >= start && steps < 4) lo--, steps++;
    while (lo + 1 < nblk && ck[lo + 1] < start && steps < 4) lo++, steps++;
    if (steps >= 4) {
        uint32_t hi = nblk;
        lo = 0;
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (ck[mid] < start) lo = mid;
            else hi = mid;
        }
    }
    const uint64_t mask = (1ULL << (2 * k)) - 1;
    const uint32_t sh = 2 * (k - 1);
    uint64_t k0 = 0, k1 = 0;
    uint32_t l = 0, len = 0;
    // t_pos of a column = checkpoint + non-insertion columns after the block's first one.  Pre-decrementing for a
    // non-insertion first column makes "every non-insertion column advances t_pos" hold from the first column on.
    uint32_t o = lo * 32;
    uint32_t tpos = ck[lo] - ((nib[o >> 1] >> 4) & 8 ? 0u : 1u);
    // skip phase, 8 columns (one aligned word) at a time: a word whose last column is still before `start` holds
    // nothing the loop below would look at
    const uint32_t *nw = reinterpret_cast<const uint32_t *>(nib);
    while (o + 8 <= n) {
        const uint32_t adv = __popc(~nw[o >> 3] & 0x88888888u);
        if (tpos + adv >= start) break;  // (a wrapped tpos = -1 only exists before the first word, whose adv >= 1)
        tpos += adv;
        o += 8;
    }
    // column loop, eight columns (one aligned word, o is a multiple of 8 here) per round from registers: the next
    // word is requested before this one is decoded, so no column waits for its own load
    uint32_t w = o < n ? nw[o >> 3] : 0;
    for (bool stop = false; o < n && !stop; o += 8) {
        const uint32_t wn = o + 8 < n ? nw[(o >> 3) + 1] : 0;
        const uint32_t cols = min(8u, n - o);
        // column c of the word in bits 4c .. 4c+3 (memory order has the first column in the HIGH nibble of each byte)
        const uint32_t r = (w & 0x0F0F0F0Fu) << 4 | (w >> 4 & 0x0F0F0F0Fu);
#pragma unroll
        for (uint32_t c = 0; c < 8; c++) {
            if (c >= cols) break;
            const uint32_t v = r >> (4 * c) & 15;
            if (!(v & 8)) tpos++;
            const uint32_t q = v & 7;
            if (tpos >= start && q != 4) {
                if (tpos <= end) {
                    if (WRITE) out[len] = code_char(q);
                    len++;
                }
                if (!WRITE && l < k) {
                    k0 = (k0 << 2 | (uint64_t)q) & mask;
                    k1 = (k1 >> 2) | (uint64_t)(3 ^ q) << sh;
                    l++;
                }
                if (tpos > end && (WRITE || l >= k)) {
                    stop = true;
                    break;
                }
            }
            if (tpos > limit) {
                stop = true;
                break;
            }
        }
        w = wn;
    }
    ScanOut so;
    so.len = len;
    so.kmer = l >= k ? yak_hash64(k0 < k1 ? k0 : k1, mask) : 0xFFFFFFFFFFFFFFFFULL;
    return so;
}
// the ref read is alignseq 0: never trimmed, stored or dropped; its columns are the contig's codes
template <bool WRITE>
__device__ __forceinline__ ScanOut scan_ref_region(const uint8_t *__restrict__ code, uint32_t L, uint32_t start,
                                                   uint32_t end, uint32_t limit, uint32_t k, uint8_t *out) {
    const uint64_t mask = (1ULL << (2 * k)) - 1;
    const uint32_t sh = 2 * (k - 1);
    uint64_t k0 = 0, k1 = 0;
    uint32_t l = 0, len = 0;
    for (uint32_t p = start; p < L; p++) {
        const uint32_t q = code[p];
        if (q != 4) {
            if (p <= end) {
                if (WRITE) out[len] = code_char(q);
                len++;
            }
            if (!WRITE && l < k) {
                k0 = (k0 << 2 | (uint64_t)q) & mask;
                k1 = (k1 >> 2) | (uint64_t)(3 ^ q) << sh;
                l++;
            }
            if (p > end && (WRITE || l >= k)) break;
        }
        if (p > limit) break;
    }
    ScanOut so;
    so.len = len;
    so.kmer = l >= k ? yak_hash64(k0 < k1 ? k0 : k1, mask) : 0xFFFFFFFFFFFFFFFFULL;
    return so;
}

// one warp per read, a lane per region the read covers: the lanes of a warp walk the same read
__global__ void __launch_bounds__(32 * kWarpsPerCta) k_pair_scan(GenoDev g, ReadsDev R, uint32_t k) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (g.cnt[C_ABORT] || i >= R.n_reads) return;
    const uint32_t np = g.rd_np[i];
    if (!np) return;
    const uint32_t j = g.rd_j[i], poff = g.rd_poff[i], limit = g.end[j] + k;
    for (uint32_t x = lane; x < np; x += 32) {
        const uint32_t reg = j + x;
        ScanOut so = scan_read_region<false>(R, i, g.start[reg], g.end[reg], limit, k, nullptr);
        g.p_len[poff + x] = so.len;
        g.p_kmer[poff + x] = so.kmer;
    }
}

/* ---------------------------------------------------------------- per-region candidate selection */

// One warp per region.  Reads are in BAM order (sorted by pos <= t_s), so the reads that can cover the region sit
// in the window pos in [start - max_span, start]; inside it a read covers region r iff j <= r <= s.
__global__ void __launch_bounds__(32 * kWarpsPerCta) k_region_select(GenoDev g, ReadsDev R,
                                                                     const uint8_t *__restrict__ blank,
                                                                     const uint8_t *__restrict__ code, uint32_t L,
                                                                     uint32_t k, uint32_t max_span,
                                                                     const uint32_t *__restrict__ first_ge, uint32_t W) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (g.cnt[C_ABORT] || r >= g.cnt[C_NREG]) return;
    const uint32_t start = g.start[r], end = g.end[r];
    const uint32_t base = r * kMaxCand;
    uint32_t ncand = 0, bytes = 0;
    if (lane == 0) {  // ref candidate (order 0); its decode limit is that of region 0 (main.rs:1467 with j = 0)
        ScanOut so = scan_ref_region<false>(code, L, start, end, g.end[0] + k, k, nullptr);
        if (so.len) {
            g.c_src[base] = 0xFFFFFFFFu;
            g.c_len[base] = so.len;
            g.c_order[base] = 0;
            g.c_kmer[base] = so.kmer;
        }
        ncand = so.len ? 1 : 0;
        bytes = so.len;
    }
    ncand = __shfl_sync(0xFFFFFFFFu, ncand, 0);
    bytes = __shfl_sync(0xFFFFFFFFu, bytes, 0);
    // a superset of the reads with pos in [start - max_span, start], from the per-stripe read index (the test below is
    // exact: a read covers region r iff j <= r <= s)
    const uint32_t lo_pos = start > max_span ? start - max_span : 0;
    const uint32_t first = first_ge[lo_pos / W], last = first_ge[start / W + 1];
    for (uint32_t i0 = first; i0 < last && ncand < kMaxCand; i0 += 32) {
        const uint32_t i = i0 + lane;
        uint32_t len = 0, pair = 0;
        if (i < last && !blank[i] && g.rd_np[i]) {
            const uint32_t j = g.rd_j[i];
            if (j <= r && r < j + g.rd_np[i]) {
                pair = g.rd_poff[i] + (r - j);
                len = g.p_len[pair];
            }
        }
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, len > 0);  // empty candidates are not pushed (main.rs:1509)
        const uint32_t slot = ncand + __popc(bal & ((1u << lane) - 1));
        if (len > 0 && slot < kMaxCand) {
            g.c_src[base + slot] = i;  // the read; k_cand_write re-decodes it over the region
            g.c_len[base + slot] = len;
            g.c_order[base + slot] = g.rd_order[i];
            g.c_kmer[base + slot] = g.p_kmer[pair];
        }
        uint32_t add = (len > 0 && slot < kMaxCand) ? len : 0;
        for (int d = 16; d > 0; d >>= 1) add += __shfl_xor_sync(0xFFFFFFFFu, add, d);
        bytes += add;
        ncand = min(ncand + __popc(bal), (uint32_t)kMaxCand);
    }
    if (lane == 0) {
        g.r_ncand[r] = ncand;
        g.r_bytes[r] = bytes;
    }
}

// One warp per region: candidate offsets inside the region's slice of the pool, then the bytes.
__global__ void __launch_bounds__(32 * kWarpsPerCta) k_cand_write(GenoDev g, ReadsDev R,
                                                                  const uint8_t *__restrict__ code, uint32_t L,
                                                                  uint32_t k) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (g.cnt[C_ABORT] || r >= g.cnt[C_NREG]) return;
    const uint32_t n = g.r_ncand[r], base = r * kMaxCand;
    const uint32_t start = g.start[r], end = g.end[r];
    uint64_t run = g.r_pool_off[r];
    for (uint32_t c0 = 0; c0 < n; c0 += 32) {
        const uint32_t c = c0 + lane;
        const uint32_t len = c < n ? g.c_len[base + c] : 0;
        uint32_t incl = len;
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += t;
        }
        const uint64_t off = run + incl - len;
        run += __shfl_sync(0xFFFFFFFFu, incl, 31);
        if (c < n) {
            g.c_off[base + c] = off;
            const uint32_t src = g.c_src[base + c];
            if (src == 0xFFFFFFFFu) {
                scan_ref_region<true>(code, L, start, end, g.end[0] + k, k, g.pool + off);
            } else {
                scan_read_region<true>(R, src, start, end, g.end[g.rd_j[src]] + k, k, g.pool + off);
            }
        }
    }
}

/* ---------------------------------------------------------------- candidate kscore */

__device__ __forceinline__ uint32_t g_bucket_of(uint64_t tag, uint32_t nb) {
    uint32_t m = (uint32_t)((tag * 0x9E3779B97F4A7C15ULL) >> 32);
    return (uint32_t)(((uint64_t)m * nb) >> 32);
}
__device__ __forceinline__ uint32_t g_probe(const TableDev &t, uint64_t h, uint32_t min_count) {
    const uint32_t sub = (uint32_t)(h & 1023);
    const uint64_t tag = h >> 10;
    uint32_t b = g_bucket_of(tag, t.nb);
    for (uint32_t step = 0; step < t.nb; step++) {
        uint64_t v[4];
        ld_bucket(t.slots + ((uint64_t)sub * t.nb + b) * kBucketSlots, v);
        bool empty = false;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if ((v[j] >> 10) == tag && v[j] != 0) {
                const uint32_t c = (uint32_t)(v[j] & 1023);
                return c >= min_count ? c : 0;
            }
            empty |= v[j] == 0;
        }
        if (empty) return 0;
        b = b + 1 == t.nb ? 0 : b + 1;
    }
    return 0;
}
// candidates no longer than k: the single pre-computed first-k k-mer (main.rs:770-774); INVALID_KMER keeps 0.
// The (rare) longer ones are queued for the warp-per-candidate kernel below.
__global__ void k_cand_kscore_short(GenoDev g, TableDev t, uint32_t min_count, uint32_t *long_count) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (g.cnt[C_ABORT] || s >= g.cnt[C_NREG] * kMaxCand) return;
    const uint32_t r = s / kMaxCand, c = s - r * kMaxCand;
    if (c >= g.r_ncand[r]) return;
    if (g.c_len[s] > t.k) {
        g.long_list[atomicAdd(long_count, 1u)] = s;
        return;
    }
    const uint64_t h = g.c_kmer[s];
    g.c_kscore[s] = h == 0xFFFFFFFFFFFFFFFFULL ? 0 : (uint16_t)g_probe(t, h, min_count);
}
// candidates longer than k: min over all their k-mers (main.rs:760-769); warps stride over the queue, k < 32 here
__global__ void __launch_bounds__(32 * kWarpsPerCta) k_cand_kscore_long(GenoDev g, TableDev t, uint32_t min_count) {
    const uint32_t lane = threadIdx.x & 31;
    if (g.cnt[C_ABORT]) return;
    const uint32_t n_long = g.cnt[C_NLONG];
    const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_long; w += nw) {
        const uint32_t s = g.long_list[w];
        const uint32_t len = g.c_len[s], k = t.k;
        const uint8_t *sq = g.pool + g.c_off[s];
        const uint64_t mask = (1ULL << (2 * k)) - 1;
        uint32_t mn = 0xFFFFFFFFu;
        for (uint32_t e = k - 1 + lane; e < len; e += 32) {
            bool ok = true;
            uint64_t f = 0, rv = 0;
            for (uint32_t x = e + 1 - k; x <= e; x++) {
                const uint32_t cd = seq_code(sq[x]);
                ok &= cd < 4;
                f = (f << 2 | cd) & mask;
                rv = (rv >> 2) | (uint64_t)(3 ^ cd) << (2 * (k - 1));
            }
            if (ok) mn = min(mn, g_probe(t, yak_hash64(f < rv ? f : rv, mask), min_count));
        }
        mn = __reduce_min_sync(0xFFFFFFFFu, mn);
        if (lane == 0) g.c_kscore[s] = mn == 0xFFFFFFFFu ? 0 : (uint16_t)mn;
    }
}

/* ---------------------------------------------------------------- genotype rules (one warp per region) */

constexpr int kRegionPool = 1536;  // bytes of candidate strings staged per region (a longer slice is read from HBM)
struct RegionSmem {
    uint32_t len[kMaxCand];
    uint32_t order[kMaxCand];
    uint64_t off[kMaxCand];
    uint16_t kscore[kMaxCand];
    uint8_t rep[kMaxCand];      // first candidate holding the same string
    uint8_t per_pos[kMaxCand];  // group size as fill_order_stat leaves it in `stats`
    uint8_t ostat[kMaxCand];    // order_stat value of the candidate's own order (0 = absent)
    uint8_t surv[kMaxCand];
    uint32_t max1_c, max1_p, max2_c, max2_p;
    const uint8_t *str0;        // candidate c's string = str0 + off[c]
    uint8_t pool[kRegionPool];
};

// Loads the candidates of region r; the region's slice of the string pool (contiguous, k_cand_write) is staged in
// shared memory so that the all-pairs string comparisons below do not go to HBM byte by byte.
// have_rep: c_rep already holds "first candidate with the same string" (k_region_hete ran on these candidates).
__device__ __forceinline__ void region_load(const GenoDev &g, uint32_t r, uint32_t n, uint32_t lane, RegionSmem &sm,
                                            bool have_rep) {
    const uint32_t base = r * kMaxCand;
    for (uint32_t c = lane; c < n; c += 32) {
        sm.len[c] = g.c_len[base + c];
        sm.order[c] = g.c_order[base + c];
        sm.off[c] = g.c_off[base + c];
        sm.kscore[c] = g.c_kscore[base + c];
        if (have_rep) sm.rep[c] = g.c_rep[base + c];
    }
    const uint64_t p0 = g.r_pool_off[r];
    const uint32_t bytes = g.r_bytes[r];
    const bool staged = bytes <= kRegionPool;
    if (staged)
        for (uint32_t x = lane; x < bytes; x += 32) sm.pool[x] = g.pool[p0 + x];
    if (lane == 0) sm.str0 = staged ? sm.pool - p0 : g.pool;
    __syncwarp();
    if (have_rep) return;
    // rep[c] = first index holding the same string (exact byte comparison)
    const uint8_t *str0 = sm.str0;
    for (uint32_t c = lane; c < n; c += 32) {
        uint32_t rp = c;
        const uint8_t *a = str0 + sm.off[c];
        const uint32_t la = sm.len[c];
        for (uint32_t d = 0; d < c; d++) {
            if (sm.len[d] != la) continue;
            const uint8_t *b = str0 + sm.off[d];
            bool eq = true;
            for (uint32_t x = 0; x < la; x++)
                if (a[x] != b[x]) {
                    eq = false;
                    break;
                }
            if (eq) {
                rp = d;
                break;
            }
        }
        sm.rep[c] = (uint8_t)rp;
    }
    __syncwarp();
}
// fill_order_stat (main.rs:813-849).  The reference walks the candidates in order; a candidate with kscore > 0 that no
// earlier one of the same string has claimed becomes the LEADER of its string: its count is the number of candidates
// from itself on that hold the string (kscore or not), and every one of those gets that count as per_pos.  All lanes
// work that out for two candidates each; lane 0 then applies the reference's max1 / max2 update to the leaders in
// ascending order (a handful per region).
__device__ __forceinline__ void region_order_stat(uint32_t n, uint32_t lane, RegionSmem &sm) {
    uint32_t lead_mask[2] = {0, 0};
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const uint32_t c = lane + 32 * h;
        uint32_t lead = 0xFFu, cnt = 0;
        if (c < n) {
            const uint32_t rc = sm.rep[c];
            for (uint32_t x = 0; x < n; x++) {
                if (sm.rep[x] != rc) continue;
                if (lead == 0xFFu && sm.kscore[x] > 0) lead = x;
                if (lead != 0xFFu) cnt++;
            }
            sm.per_pos[c] = (lead != 0xFFu && lead <= c) ? (uint8_t)cnt : 0;
            sm.ostat[c] = lead == c ? (uint8_t)cnt : 0;
        }
        lead_mask[h] = __ballot_sync(0xFFFFFFFFu, c < n && lead == c);
    }
    __syncwarp();
    if (lane == 0) {
        uint32_t max1_c = 0, max1_p = 0, max2_c = 0, max2_p = 0;
        for (int h = 0; h < 2; h++)
            for (uint32_t mk = lead_mask[h]; mk; mk &= mk - 1) {
                const uint32_t p1 = (uint32_t)__ffs(mk) - 1 + 32 * h, c = sm.ostat[p1];
                if (c > max1_c || (c == max1_c && sm.order[p1] == 0)) {
                    max2_c = max1_c;
                    max2_p = max1_p;
                    max1_c = c;
                    max1_p = p1;
                } else if (max1_p == max2_p || c > max2_c) {
                    max2_c = c;
                    max2_p = p1;
                }
            }
        sm.max1_c = max1_c;
        sm.max1_p = max1_p;
        sm.max2_c = max2_c;
        sm.max2_p = max2_p;
    }
    __syncwarp();
}
__device__ __forceinline__ uint32_t min_count_for(uint32_t c) { return c >= 9 ? 3 : (c >= 6 ? 2 : 1); }  // main.rs:803
// is_valid_snp (main.rs:780-801)
__device__ __forceinline__ bool hp_differ(const uint8_t *a, uint32_t na, const uint8_t *b, uint32_t nb) {
    uint32_t i = 0, j = 0;
    while (i < na && j < nb) {
        if (a[i] != b[j]) return true;
        while (i + 1 < na && a[i] == a[i + 1]) i++;
        while (j + 1 < nb && b[j] == b[j + 1]) j++;
        i++;
        j++;
    }
    return false;
}

// mark_hete_lqseqs (main.rs:916-946) + the number of agreement edges the region will emit
__global__ void __launch_bounds__(32 * kWarpsPerCta) k_region_hete(GenoDev g) {
    __shared__ RegionSmem smem[kWarpsPerCta];
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (g.cnt[C_ABORT] || r >= g.cnt[C_NREG]) return;
    RegionSmem &sm = smem[threadIdx.x >> 5];
    const uint32_t n = g.r_ncand[r];
    region_load(g, r, n, lane, sm, false);
    if (n) region_order_stat(n, lane, sm);
    if (lane == 0) {
        uint8_t lable = 0;
        uint32_t nedge = 0;
        if (n) {
            const uint32_t min_c = min_count_for(n);
            const uint32_t a = sm.max1_p, b = sm.max2_p;
            if (sm.max2_c >= min_c && (sm.len[a] == sm.len[b] || (n >= 6 && sm.max2_c >= sm.max1_c / 2)) &&
                hp_differ(sm.str0 + sm.off[a], sm.len[a], sm.str0 + sm.off[b], sm.len[b])) {
                lable = 0x40;
                uint32_t m = 0;
                for (uint32_t p = 0; p < n; p++) {
                    if (sm.kscore[p] > 0 && sm.per_pos[p] < min_c) {
                        sm.kscore[p] = 0;
                        g.c_kscore[r * kMaxCand + p] = 0;
                    }
                    m += sm.kscore[p] > 0;
                }
                nedge = m * (m - 1) / 2;
            }
        }
        g.r_lable[r] = lable;
        g.r_nedge[r] = nedge;
    }
    // rep is needed again by k_edges_accum
    for (uint32_t c = lane; c < n; c += 32) g.c_rep[r * kMaxCand + c] = sm.rep[c];
}

// all pairs (i < j) of supported candidates of a heterozygous region (main.rs:953-992): +1 when the strings agree,
// 2^32 - 1 (a "differs" count in the high half, -1 in the low half) when they do not.
// Reads are ordered by position, so the partners y > x of read x lie in the index window (x, x + W_x] (W_x = reads that
// start before x ends, computed once on the host): the pair (x, y) owns slot pair_off[x] + (y - x - 1) of a dense
// accumulator.  No pair list, no sort: every observation is one 64-bit atomic add, and the non-zero slots read in slot
// order are the reduced pair records in (x, y) order.
__global__ void __launch_bounds__(32 * kWarpsPerCta) k_edges_accum(GenoDev g, const uint64_t *__restrict__ pair_off,
                                                                   unsigned long long *__restrict__ acc, uint32_t *err) {
    __shared__ uint8_t s_valid[kWarpsPerCta][kMaxCand];
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (g.cnt[C_ABORT] || r >= g.cnt[C_NREG]) return;
    if (!g.r_nedge[r]) return;
    uint8_t *valid = s_valid[threadIdx.x >> 5];
    const uint32_t n = g.r_ncand[r], base = r * kMaxCand;
    uint32_t m = 0;
    if (lane == 0) {
        for (uint32_t p = 0; p < n; p++)
            if (g.c_kscore[base + p] > 0) valid[m++] = (uint8_t)p;
    }
    m = __shfl_sync(0xFFFFFFFFu, m, 0);
    __syncwarp();
    for (uint32_t a = 0; a + 1 < m; a++) {
        const uint32_t pa = valid[a];
        const uint32_t oa = g.c_order[base + pa], ra = g.c_rep[base + pa];
        for (uint32_t b = a + 1 + lane; b < m; b += 32) {
            const uint32_t pb = valid[b];
            const uint32_t ob = g.c_order[base + pb];
            const bool same = g.c_rep[base + pb] == ra;
            const uint32_t x = min(oa, ob), y = max(oa, ob);
            const uint64_t slot = pair_off[x] + (y - x - 1);
            if (slot >= pair_off[x + 1]) {
                atomicExch(err, 4u);
                continue;
            }
            atomicAdd(acc + slot, same ? 1ULL : 0xFFFFFFFFULL);
        }
    }
}
// W[x] = number of alignseqs y > x that start (record pos) at or before x's last column; x = 0 (ref) pairs with all
__global__ void k_pair_windows(const uint32_t *__restrict__ as_pos, const uint32_t *__restrict__ as_te, uint32_t na,
                               uint32_t *__restrict__ W) {
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a > na) return;
    if (a == na) {
        W[a] = 0;
        return;
    }
    if (a == 0) {
        W[0] = na - 1;
        return;
    }
    const uint32_t te = as_te[a];
    uint32_t lo = a + 1, hi = na;  // first y in (a, na) with as_pos[y] > te
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (as_pos[mid] <= te) lo = mid + 1;
        else hi = mid;
    }
    W[a] = lo - (a + 1);
}
void geno_pair_windows(const uint32_t *d_as_pos, const uint32_t *d_as_te, uint32_t na, uint32_t *d_W, cudaStream_t s) {
    NP2_K(k_pair_windows)<<<cdiv(na + 1, 256), 256, 0, s>>>(d_as_pos, d_as_te, na, d_W);
}
void geno_pair_window_offsets(const uint32_t *d_W, uint64_t *d_pair_off, uint32_t na, ScanPool &pool, cudaStream_t s) {
    ScanOffsets<uint32_t, uint64_t> f;
    f.in = d_W;
    f.out = d_pair_off;
    f.c_slot = nullptr;
    f.q_slot = nullptr;
    f.cap = ~0ULL;
    f.abort = nullptr;
    scan_launch(f, nullptr, 0, na, pool, s);
}
// selected slots -> pair records (key = x << 32 | y, value) in (x, y) order
__global__ void k_edges_finish(const uint32_t *__restrict__ sel, const uint32_t *__restrict__ cnt,
                               const uint64_t *__restrict__ pair_off, uint32_t n_ids,
                               const unsigned long long *__restrict__ acc, uint64_t *__restrict__ key,
                               long long *__restrict__ val) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (cnt[C_ABORT] || i >= cnt[C_NU]) return;
    const uint32_t slot = sel[i];
    uint32_t lo = 0, hi = n_ids;  // largest x with pair_off[x] <= slot
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (pair_off[mid] <= slot) lo = mid;
        else hi = mid;
    }
    key[i] = (uint64_t)lo << 32 | (lo + 1 + (uint32_t)(slot - pair_off[lo]));
    val[i] = (long long)acc[slot];
}

// fill_seed_lqseqs (main.rs:862-914) with retain_sort_seqs (714-726)
__global__ void __launch_bounds__(32 * kWarpsPerCta) k_region_seed(GenoDev g, int32_t max_indel_len, uint32_t *err,
                                                                   int have_rep) {
    __shared__ RegionSmem smem[kWarpsPerCta];
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (g.cnt[C_ABORT] || r >= g.cnt[C_NREG]) return;
    RegionSmem &sm = smem[threadIdx.x >> 5];
    const uint32_t n = g.r_ncand[r];
    region_load(g, r, n, lane, sm, have_rep != 0);
    if (n == 0 || sm.order[0] != 0) {  // reference would panic: no candidate / "the first lqseq is not ref."
        if (lane == 0) atomicExch(err, n == 0 ? 1u : 2u);
        return;
    }
    region_order_stat(n, lane, sm);
    if (lane != 0) return;
    uint32_t seed = sm.max1_p;
    const uint32_t min_c = min_count_for(n), max1_c = sm.max1_c, max1_p = sm.max1_p;
    if (sm.ostat[0]) {  // keep the reference allele when it has support (main.rs:876-890)
        if (sm.ostat[0] > 1 && sm.ostat[0] < min_c) sm.ostat[0] = (uint8_t)min_c;
    } else {
        uint32_t c = 0;
        for (uint32_t x = 0; x < n; x++) c += sm.rep[x] == 0;
        if (c > 1) sm.ostat[0] = (uint8_t)min_c;
    }
    bool no_dup = true;  // no_dupseq_lqseq main.rs:851-860: no two equal strings among candidates 1..n-1
    for (uint32_t p1 = 1; p1 < n && no_dup; p1++)
        for (uint32_t p2 = p1 + 1; p2 < n; p2++)
            if (sm.rep[p1] == sm.rep[p2]) {
                no_dup = false;
                break;
            }
    if (max1_p != 0 && max1_c < min_c && (max1_c > 1 || no_dup)) {
        sm.ostat[max1_p] = (uint8_t)min_c;
        sm.ostat[0] = (uint8_t)min_c;
    } else if (max1_c < min_c) {
        sm.ostat[0] = (uint8_t)min_c;
    }
    // stable sort by order_stat descending, keep those >= min_c
    uint32_t ns = 0;
    for (uint32_t p = 0; p < n; p++) {
        if (sm.ostat[p] < min_c) continue;  // everything below min_c sorts after the cut anyway
        uint32_t q = ns++;
        while (q > 0 && sm.ostat[sm.surv[q - 1]] < sm.ostat[p]) {
            sm.surv[q] = sm.surv[q - 1];
            q--;
        }
        sm.surv[q] = (uint8_t)p;
    }
    if (ns == 0) {
        atomicExch(err, 3u);
        return;
    }
    uint8_t lable = 0x80 | 0x20;
    const int32_t dl = (int32_t)sm.len[seed] - (int32_t)sm.len[sm.surv[0]];
    const bool skip_long = (dl < 0 ? -dl : dl) > max_indel_len;
    if (ns <= 1 || skip_long) {
        seed = sm.surv[0];
        lable ^= 0x20;
        ns = 0;
    }
    g.r_lable[r] = lable;
    g.r_seed_off[r] = sm.off[seed];
    g.r_seed_len[r] = sm.len[seed];
    g.r_nsurv[r] = ns;
    for (uint32_t q = 0; q < ns; q++) g.r_surv[r * kMaxCand + q] = sm.surv[q];
}

/* ---------------------------------------------------------------- launch wrappers */

namespace {
inline uint32_t region_grid(uint32_t cap_reg) { return cdiv((uint64_t)cap_reg * 32, 32 * kWarpsPerCta); }
}  // namespace
void geno_read_cursor(GenoDev g, const ReadsDev &R, const uint8_t *d_blank, cudaStream_t s) {
    if (R.n_reads) NP2_K(k_read_cursor)<<<cdiv(R.n_reads, 256), 256, 0, s>>>(g, R, d_blank);
}
void geno_cursor_min(uint32_t *d_rd_s, uint32_t n_reads, const uint32_t *d_abort, ScanPool &pool, cudaStream_t s) {
    if (n_reads) scan_launch(ScanInclusiveMinU32{d_rd_s}, nullptr, 0, n_reads, pool, s, d_abort);
}
void geno_read_ranges(GenoDev g, const ReadsDev &R, const uint8_t *d_blank, uint32_t k, cudaStream_t s) {
    if (R.n_reads) NP2_K(k_read_ranges)<<<cdiv(R.n_reads, 256), 256, 0, s>>>(g, R, d_blank, k);
}
void geno_pair_offsets(GenoDev g, uint32_t n_reads, uint32_t cap_pairs, CountsDev cd, ScanPool &pool, cudaStream_t s) {
    ScanOffsets<uint32_t, uint32_t> f;
    f.in = g.rd_np;
    f.out = g.rd_poff;
    f.c_slot = cd.c + C_NPAIRS;
    f.q_slot = nullptr;
    f.cap = cap_pairs;
    f.abort = cd.c + C_ABORT;
    scan_launch(f, nullptr, 0, n_reads, pool, s, cd.c + C_ABORT);
}
void geno_pair_scan(GenoDev g, const ReadsDev &R, uint32_t k, uint32_t cap_pairs, cudaStream_t s) {
    if (cap_pairs && R.n_reads)
        NP2_K(k_pair_scan)<<<cdiv((uint64_t)R.n_reads * 32, 32 * kWarpsPerCta), 32 * kWarpsPerCta, 0, s>>>(g, R, k);
}
void geno_region_select(GenoDev g, const ReadsDev &R, const uint8_t *d_blank, const uint8_t *d_code, uint32_t L,
                        uint32_t k, uint32_t max_span, const uint32_t *d_first_ge, uint32_t stripe_width, uint32_t cap_reg,
                        cudaStream_t s) {
    if (cap_reg)
        NP2_K(k_region_select)<<<region_grid(cap_reg), 32 * kWarpsPerCta, 0, s>>>(g, R, d_blank, d_code, L, k, max_span, d_first_ge,
                                                                                  stripe_width);
}
void geno_pool_offsets(GenoDev g, uint32_t cap_reg, unsigned long long cap_pool, CountsDev cd, ScanPool &pool,
                       cudaStream_t s) {
    ScanOffsets<uint32_t, uint64_t> f;  // 64-bit offsets out of 32-bit per-region byte counts
    f.in = g.r_bytes;
    f.out = g.r_pool_off;
    f.c_slot = nullptr;
    f.q_slot = cd.q + Q_POOL;
    f.cap = cap_pool;
    f.abort = cd.c + C_ABORT;
    scan_launch(f, cd.c + C_NREG, 0, cap_reg, pool, s, cd.c + C_ABORT);
}
void geno_cand_write(GenoDev g, const ReadsDev &R, const uint8_t *d_code, uint32_t L, uint32_t k, uint32_t cap_reg,
                     cudaStream_t s) {
    if (cap_reg) NP2_K(k_cand_write)<<<region_grid(cap_reg), 32 * kWarpsPerCta, 0, s>>>(g, R, d_code, L, k);
}
void geno_cand_kscore(GenoDev g, const TableDev &t, uint32_t min_count, uint32_t cap_reg, CountsDev cd, cudaStream_t s) {
    if (!cap_reg) return;
    const uint64_t slots = (uint64_t)cap_reg * kMaxCand;
    NP2_K(k_cand_kscore_short)<<<cdiv(slots, 256), 256, 0, s>>>(g, t, min_count, cd.c + C_NLONG);
    NP2_K(k_cand_kscore_long)<<<148 * 4, 32 * kWarpsPerCta, 0, s>>>(g, t, min_count);
}
void geno_region_hete(GenoDev g, uint32_t cap_reg, cudaStream_t s) {
    if (cap_reg) NP2_K(k_region_hete)<<<region_grid(cap_reg), 32 * kWarpsPerCta, 0, s>>>(g);
}
void geno_edge_offsets(GenoDev g, uint32_t cap_reg, CountsDev cd, ScanPool &pool, cudaStream_t s) {
    ScanOffsets<uint32_t, uint64_t> f;
    f.in = g.r_nedge;
    f.out = g.r_edge_off;
    f.c_slot = nullptr;
    f.q_slot = cd.q + Q_EDGES;
    f.cap = ~0ULL;
    f.abort = nullptr;
    scan_launch(f, cd.c + C_NREG, 0, cap_reg, pool, s, cd.c + C_ABORT);
}
void geno_edges_accum(GenoDev g, const uint64_t *d_pair_off, unsigned long long *d_acc, uint32_t cap_reg, CountsDev cd,
                      cudaStream_t s) {
    if (cap_reg) NP2_K(k_edges_accum)<<<region_grid(cap_reg), 32 * kWarpsPerCta, 0, s>>>(g, d_pair_off, d_acc, cd.c + C_PERR);
}
// non-zero slots in slot order
void geno_edges_select(const unsigned long long *d_acc, uint32_t n_slots, uint32_t *d_sel, uint32_t cap_nu, CountsDev cd,
                       ScanPool &pool, cudaStream_t s) {
    ScanSelect<PredNonZeroU64> f;
    f.pred = PredNonZeroU64{d_acc};
    f.out = d_sel;
    f.count = cd.c + C_NU;
    f.cap = cap_nu;
    f.abort = cd.c + C_ABORT;
    scan_launch(f, nullptr, 0, n_slots, pool, s, cd.c + C_ABORT);
}
void geno_edges_finish(const uint32_t *d_sel, uint32_t cap_nu, const uint64_t *d_pair_off, uint32_t n_ids,
                       const unsigned long long *d_acc, uint64_t *d_key, long long *d_val, CountsDev cd, cudaStream_t s) {
    if (cap_nu) NP2_K(k_edges_finish)<<<cdiv(cap_nu, 256), 256, 0, s>>>(d_sel, cd.c, d_pair_off, n_ids, d_acc, d_key, d_val);
}
/* ---------------------------------------------------------------- level 0 of the phasing graph (np2_phase.cpp)
 * From the pair accumulator (slot (a, b), a = 0 is the ref read) to what the host Louvain starts from: per-read flags,
 * and the adjacency in CSR form with the `dif <= -3` override applied and the reads that disagree with the ref read
 * removed (main.rs:972-1010). */
/* Everything is read from the dense accumulator, whose slot order IS the (x, y) order.  A read v has its later partners in its own window (slots pair_off[v] .. pair_off[v + 1]) and its
 * earlier partners u wherever u's window reaches v; those u start at most max_span before v (reads are in position
 * order), so they are found by one binary search and a short scan.  Both halves come out in ascending partner order,
 * which is the order the host Louvain needs, so no sort is involved. */
__device__ __forceinline__ float pair_weight(long long v) {  // main.rs:972-992: `dif <= -3` override
    const long long ndif = (v + (1LL << 31)) >> 32;
    return ndif >= 3 ? -(float)ndif : (float)(v - (ndif << 32));
}
__global__ void k_phase_ref_acc(const unsigned long long *__restrict__ acc, const uint64_t *__restrict__ pair_off, uint32_t n,
                                const uint32_t *__restrict__ cnt, PhaseDev p, int asref, int use_all) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (cnt[C_ABORT] || b >= n) return;
    const long long v = (long long)acc[pair_off[0] + b - 1];  // the ref read's window holds every read
    if (!v) return;
    const long long ndif = (v + (1LL << 31)) >> 32;
    if (asref) {
        p.ref_w[b] = (float)(v - (ndif << 32));
        p.in_ref[b] = 1;
    }
    if (ndif > 0 && !use_all) p.bad_v[b] = 1;
}
template <bool FILL>
__global__ void __launch_bounds__(32 * kWarpsPerCta) k_phase_adj(const unsigned long long *__restrict__ acc,
                                                                 const uint64_t *__restrict__ pair_off,
                                                                 const uint32_t *__restrict__ as_pos, uint32_t n,
                                                                 uint32_t max_span, uint32_t *__restrict__ cnt, PhaseDev p,
                                                                 int use_all, uint32_t *__restrict__ deg,
                                                                 const uint32_t *__restrict__ aoff,
                                                                 uint32_t *__restrict__ ato, float *__restrict__ aw,
                                                                 uint32_t cap_dir) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (cnt[C_ABORT] || v >= n) return;
    const uint64_t s0 = pair_off[v], s1 = pair_off[v + 1];
    if (v == 0) {  // the ref read is no vertex; its non-zero slots only count as pair records
        if (!FILL) {
            uint32_t nz = 0;
            for (uint64_t sl = s0 + lane; sl < s1; sl += 32) nz += acc[sl] != 0;
            for (int d = 16; d > 0; d >>= 1) nz += __shfl_xor_sync(0xFFFFFFFFu, nz, d);
            if (lane == 0) {
                if (nz) atomicAdd(cnt + C_NU, nz);
                deg[0] = 0;
                deg[n] = 0;
            }
        }
        return;
    }
    const bool bad_self = !use_all && p.bad_v[v];
    uint32_t w_out = FILL ? aoff[v] : 0, n_edge = 0, nz_fwd = 0;
    bool any = false;
    // earlier partners: u in [lo, v) with as_pos[u] + max_span >= as_pos[v] (u >= 1)
    const uint32_t pv = as_pos[v], want = pv > max_span ? pv - max_span : 0;
    uint32_t lo = 1, hi = v;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (as_pos[mid] < want) lo = mid + 1;
        else hi = mid;
    }
    for (uint32_t u0 = lo; u0 < v; u0 += 32) {
        const uint32_t u = u0 + lane;
        unsigned long long val = 0;
        if (u < v) {
            const uint64_t sl = pair_off[u] + (v - u - 1);
            if (sl < pair_off[u + 1]) val = acc[sl];
        }
        const bool edge = val != 0 && !bad_self && !(!use_all && p.bad_v[u]);
        any |= val != 0;
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, edge);
        if (FILL && edge) {
            const uint32_t at = w_out + __popc(bal & ((1u << lane) - 1));
            if (at < cap_dir) {
                ato[at] = u;
                aw[at] = pair_weight((long long)val);
            }
        }
        w_out += __popc(bal);
        n_edge += __popc(bal);
    }
    // later partners: v's own window, y = v + 1 + (slot - s0)
    for (uint64_t b0 = s0; b0 < s1; b0 += 32) {
        const uint64_t sl = b0 + lane;
        const unsigned long long val = sl < s1 ? acc[sl] : 0;
        const uint32_t y = v + 1 + (uint32_t)(sl - s0);
        const bool edge = val != 0 && !bad_self && !(!use_all && p.bad_v[y < n ? y : v]);
        any |= val != 0;
        nz_fwd += val != 0;
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, edge);
        if (FILL && edge) {
            const uint32_t at = w_out + __popc(bal & ((1u << lane) - 1));
            if (at < cap_dir) {
                ato[at] = y;
                aw[at] = pair_weight((long long)val);
            }
        }
        w_out += __popc(bal);
        n_edge += __popc(bal);
    }
    if (!FILL) {
        any = __any_sync(0xFFFFFFFFu, any);
        for (int d = 16; d > 0; d >>= 1) nz_fwd += __shfl_xor_sync(0xFFFFFFFFu, nz_fwd, d);
        if (lane == 0) {
            deg[v] = n_edge;
            if (any && !bad_self) p.has[v] = 1;
            if (nz_fwd) atomicAdd(cnt + C_NU, nz_fwd);
        }
    }
}
void phase_ref_acc(const unsigned long long *d_acc, const uint64_t *d_pair_off, uint32_t n, CountsDev cd, PhaseDev p,
                   bool asref, bool use_all, cudaStream_t s) {
    if (n > 1) NP2_K(k_phase_ref_acc)<<<cdiv(n - 1, 256), 256, 0, s>>>(d_acc, d_pair_off, n, cd.c, p, asref, use_all);
}
void phase_adj_count(const unsigned long long *d_acc, const uint64_t *d_pair_off, const uint32_t *d_as_pos, uint32_t n,
                     uint32_t max_span, CountsDev cd, PhaseDev p, bool use_all, uint32_t *d_deg, cudaStream_t s) {
    if (n)
        NP2_K(k_phase_adj<false>)<<<cdiv((uint64_t)n * 32, 32 * kWarpsPerCta), 32 * kWarpsPerCta, 0, s>>>(
            d_acc, d_pair_off, d_as_pos, n, max_span, cd.c, p, use_all, d_deg, nullptr, nullptr, nullptr, 0);
}
void phase_adj_offsets(const uint32_t *d_deg, uint32_t *d_aoff, uint32_t n, uint32_t cap_dir, CountsDev cd, ScanPool &pool,
                       cudaStream_t s) {
    ScanOffsets<uint32_t, uint32_t> f;
    f.in = d_deg;
    f.out = d_aoff;
    f.c_slot = cd.c + C_NDIR;
    f.q_slot = nullptr;
    f.cap = cap_dir;
    f.abort = cd.c + C_ABORT;
    scan_launch(f, nullptr, 0, n, pool, s, cd.c + C_ABORT);
}
void phase_adj_fill(const unsigned long long *d_acc, const uint64_t *d_pair_off, const uint32_t *d_as_pos, uint32_t n,
                    uint32_t max_span, CountsDev cd, PhaseDev p, bool use_all, const uint32_t *d_aoff, uint32_t *d_ato,
                    float *d_aw, uint32_t cap_dir, cudaStream_t s) {
    if (n)
        NP2_K(k_phase_adj<true>)<<<cdiv((uint64_t)n * 32, 32 * kWarpsPerCta), 32 * kWarpsPerCta, 0, s>>>(
            d_acc, d_pair_off, d_as_pos, n, max_span, cd.c, p, use_all, nullptr, d_aoff, d_ato, d_aw, cap_dir);
}
void geno_region_seed(GenoDev g, int32_t max_indel_len, uint32_t cap_reg, CountsDev cd, bool have_rep, cudaStream_t s) {
    if (cap_reg)
        NP2_K(k_region_seed)<<<region_grid(cap_reg), 32 * kWarpsPerCta, 0, s>>>(g, max_indel_len, cd.c + C_GERR, have_rep ? 1 : 0);
}

}  // namespace np2
