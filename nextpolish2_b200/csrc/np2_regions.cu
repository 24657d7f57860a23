// np2_regions.cu — LQ-region detection and final consensus assembly on the device.
//
// The reference finds LQ regions with a sequential state machine while it backtracks (main.rs:1586-1625).  Only the
// sparse "events" (bases with qv < 95 or coverage < 2) change its state, and the three sequential dependencies it
// has are all local, so it parallelises exactly:
//   1. whether an open LQ run closes after a low-qv event depends only on the consensus between that event and the
//      next one (first HQ base >= 5 later whose two predecessors differ in pos and base)      -> k_event_close
//   2. the run's first event is found by walking back to the previous closing / low-coverage event -> k_region_make
//   3. a new region is merged into the previous one iff its end reaches the previous candidate's start, which is
//      a property of adjacent candidates only (main.rs:1613-1614)                             -> k_region_heads/out
// Indices: "rp" is the reference's index into the reversed consensus (rp = N - 1 - i).
#include "np2_kernels.cuh"

namespace np2 {

namespace {
inline uint32_t cdiv(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }
constexpr uint32_t kNone = 0xFFFFFFFFu;
}  // namespace

// events[t] ascending consensus index => descending rp.  The next event in the reference's scan order is t - 1.
__global__ void k_event_close(RegionDev d) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (d.cnt[C_ABORT] || t >= d.cnt[C_NEV]) return;
    const uint32_t N = d.cnt[C_N];
    const uint32_t idx = d.events[t];
    const uint32_t f = d.cflags[idx];
    uint32_t close_rp = kNone;
    uint8_t boundary = 0;
    if (f & 2) {
        boundary = 1;  // coverage < 2 drops the open run (main.rs:1586-1588)
    } else {
        const uint32_t rp = N - 1 - idx;
        const uint32_t nx = t > 0 ? N - 1 - d.events[t - 1] : N;
        // HQ bases rp' in [rp + 5, nx): closes at the first one whose predecessors differ in pos and base
        for (uint64_t x = (uint64_t)rp + 5; x < nx; x++) {
            const uint32_t i1 = N - 1 - (uint32_t)(x - 1), i2 = N - 1 - (uint32_t)(x - 2);
            if (d.cpos[i1] != d.cpos[i2] && d.cbase[i1] != d.cbase[i2]) {
                close_rp = (uint32_t)x;
                boundary = 1;
                break;
            }
        }
    }
    d.ev_close[t] = close_rp;
    d.ev_boundary[t] = boundary;
    d.ev_closes[t] = close_rp != kNone;
}

// one thread per closing event (candidate region); c_t[c] = event index, ascending
__global__ void k_region_make(RegionDev d) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (d.cnt[C_ABORT] || c >= d.cnt[C_NCAND]) return;
    const uint32_t N = d.cnt[C_N], n_ev = d.cnt[C_NEV];
    const uint32_t t = d.c_t[c];
    uint32_t lq_e = d.ev_close[t] - 2;  // main.rs:1600
    // first low-qv event of the run: walk towards smaller rp (larger t) until the previous boundary
    uint32_t tf = t;
    while (tf + 1 < n_ev && !d.ev_boundary[tf + 1]) tf++;
    uint32_t lq_s = N - 1 - d.events[tf];
    lq_s = lq_s > 2 ? lq_s - 2 : 1;  // main.rs:1601-1605
    auto P = [&](uint32_t rp) { return d.cpos[N - 1 - rp]; };
    auto B = [&](uint32_t rp) { return d.cbase[N - 1 - rp]; };
    while (lq_s > 1 && (P(lq_s - 1) == P(lq_s) || B(lq_s - 1) == B(lq_s))) lq_s--;  // main.rs:1606-1611
    const uint32_t start_pos = P(lq_e), end_pos = P(lq_s);
    uint32_t a = N - 1 - lq_e, b = N - 1 - lq_s + 1;
    while (a > 0 && d.cpos[a - 1] == start_pos) a--;
    while (b < N && d.cpos[b] <= end_pos) b++;
    d.c_start[c] = start_pos;
    d.c_end[c] = end_pos;
    d.c_a[c] = a;
    d.c_b[c] = b;
}
// candidate c (larger c = earlier in the reference's scan) is merged into c + 1 when its end reaches that start
__global__ void k_region_heads(RegionDev d) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n_cand = d.cnt[C_NCAND];
    if (d.cnt[C_ABORT] || c >= n_cand) return;
    const bool merged = c + 1 < n_cand && d.c_end[c] >= d.c_start[c + 1];
    d.c_head[c] = merged ? 0 : 1;
}
// heads in ascending c with exclusive rank hr; the reference's region order is descending c
__global__ void k_region_out(RegionDev d) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n_cand = d.cnt[C_NCAND], n_heads = d.cnt[C_NREG];
    if (d.cnt[C_ABORT] || c >= n_cand || !d.c_head[c]) return;
    uint32_t lo = c;
    while (lo > 0 && !d.c_head[lo - 1]) lo--;  // candidates lo..c-1 were merged into c
    const uint32_t r = n_heads - 1 - d.c_hrank[c];
    d.r_start[r] = d.c_start[lo];
    d.r_a[r] = d.c_a[lo];
    d.r_end[r] = d.c_end[c];
    d.r_b[r] = d.c_b[c];
}

void regions_event_close(RegionDev d, uint32_t cap_ev, cudaStream_t s) {
    if (cap_ev) NP2_K(k_event_close)<<<cdiv(cap_ev, 128), 128, 0, s>>>(d);
}
void regions_cand_select(RegionDev d, uint32_t cap_ev, uint32_t cap_cand, CountsDev cd, ScanPool &pool, cudaStream_t s) {
    ScanSelect<PredFlagU8> f;
    f.pred = PredFlagU8{d.ev_closes};
    f.out = d.c_t;
    f.count = cd.c + C_NCAND;
    f.cap = cap_cand;
    f.abort = cd.c + C_ABORT;
    scan_launch(f, cd.c + C_NEV, 0, cap_ev, pool, s, cd.c + C_ABORT);
}
void regions_make(RegionDev d, uint32_t cap_cand, cudaStream_t s) {
    if (!cap_cand) return;
    NP2_K(k_region_make)<<<cdiv(cap_cand, 128), 128, 0, s>>>(d);
    NP2_K(k_region_heads)<<<cdiv(cap_cand, 128), 128, 0, s>>>(d);
}
void regions_rank(RegionDev d, uint32_t cap_cand, uint32_t cap_reg, CountsDev cd, ScanPool &pool, cudaStream_t s) {
    ScanOffsets<uint32_t, uint32_t> f;
    f.in = d.c_head;
    f.out = d.c_hrank;
    f.c_slot = cd.c + C_NREG;
    f.q_slot = nullptr;
    f.cap = cap_reg;
    f.abort = cd.c + C_ABORT;
    scan_launch(f, cd.c + C_NCAND, 0, cap_cand, pool, s, cd.c + C_ABORT);
}
void regions_out(RegionDev d, uint32_t cap_cand, cudaStream_t s) {
    if (cap_cand) NP2_K(k_region_out)<<<cdiv(cap_cand, 128), 128, 0, s>>>(d);
}

/* ---------------------------------------------------------------- seeds / survivors / assembly */

// per region in ASCENDING position q = nreg - 1 - r: length change of the patch and seed length
__global__ void k_patch_sizes(AssembleDev a) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nreg = a.cnt[C_NREG];
    if (a.cnt[C_ABORT] || q >= nreg) return;
    const uint32_t r = nreg - 1 - q;
    a.q_delta[q] = (long long)a.r_seed_len[r] - (long long)(a.r_b[r] - a.r_a[r]);
    a.q_seedlen[q] = (!a.near || a.near[r]) ? a.r_seed_len[r] : 0;
}
// compact copy of every region's seed string (for the host's flank extraction), q order
__global__ void k_seed_gather(AssembleDev a, uint8_t *__restrict__ out) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nreg = a.cnt[C_NREG];
    if (a.cnt[C_ABORT] || q >= nreg) return;
    const uint32_t r = nreg - 1 - q;
    if (a.near && !a.near[r]) return;
    const uint8_t *src = a.pool + a.r_seed_off[r];
    uint8_t *dst = out + a.q_seedoff[q];
    for (uint32_t x = lane; x < a.r_seed_len[r]; x += 32) dst[x] = src[x];
}
// survivors of the regions that stay RECH: (order, len) entries + strings, in retain_sort_seqs order
__global__ void k_rech_sizes(GenoDev g, uint32_t *__restrict__ bytes) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (g.cnt[C_ABORT] || r >= g.cnt[C_NREG]) return;
    uint32_t b = 0;
    for (uint32_t x = 0; x < g.r_nsurv[r]; x++) b += g.c_len[r * kMaxCand + g.r_surv[r * kMaxCand + x]];
    bytes[r] = b;
}
__global__ void k_rech_gather(GenoDev g, const uint32_t *__restrict__ ent_off, const uint64_t *__restrict__ byte_off,
                              uint32_t *__restrict__ ent_order, uint32_t *__restrict__ ent_len,
                              uint64_t *__restrict__ ent_pool_off, uint8_t *__restrict__ out) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (g.cnt[C_ABORT] || r >= g.cnt[C_NREG]) return;
    uint64_t w = byte_off[r];
    for (uint32_t x = 0; x < g.r_nsurv[r]; x++) {
        const uint32_t sl = r * kMaxCand + g.r_surv[r * kMaxCand + x];
        const uint32_t len = g.c_len[sl];
        ent_order[ent_off[r] + x] = g.c_order[sl];
        ent_len[ent_off[r] + x] = len;
        ent_pool_off[ent_off[r] + x] = g.c_off[sl];  // offset in the device pool (for the final seed override)
        const uint8_t *src = g.pool + g.c_off[sl];
        for (uint32_t i = 0; i < len; i++) out[w + i] = src[i];
        w += len;
    }
}
// final consensus = DP consensus with [a, b) of every region replaced by its seed (main.rs:1027-1058).
// One warp per region q (ascending position) copies the DP stretch before it and its seed; warp nreg copies the tail.
__global__ void __launch_bounds__(128) k_assemble(AssembleDev a, uint8_t *__restrict__ out) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nreg = a.cnt[C_NREG];
    if (a.cnt[C_ABORT] || q > nreg) return;
    const long long shift = a.q_shift[q];  // sum of deltas of the regions before q
    const uint32_t lo = q ? a.r_b[nreg - q] : 0;              // region q - 1 is r = nreg - q
    const uint32_t hi = q < nreg ? a.r_a[nreg - 1 - q] : a.cnt[C_N];
    for (uint32_t i = lo + lane; i < hi; i += 32) out[(long long)i + shift] = a.cbase[i];
    if (q < nreg) {
        const uint32_t r = nreg - 1 - q;
        const uint8_t *src = a.pool + a.r_seed_off[r];
        uint8_t *dst = out + ((long long)hi + shift);
        for (uint32_t x = lane; x < a.r_seed_len[r]; x += 32) dst[x] = src[x];
    }
}

// out[off[i] .. off[i+1]) = src[lo[i] ..): one warp per range (the DP-base windows around RECH regions)
__global__ void __launch_bounds__(128) k_gather_ranges(const uint8_t *__restrict__ src, const uint32_t *__restrict__ lo,
                                                       const uint64_t *__restrict__ off, const uint32_t *__restrict__ d_n,
                                                       uint32_t cap, uint8_t *__restrict__ out) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n = d_n ? min(*d_n, cap) : cap;
    if (i >= n) return;
    const uint64_t o = off[i];
    const uint32_t len = (uint32_t)(off[i + 1] - o);
    const uint8_t *sp = src + lo[i];
    for (uint32_t x = lane; x < len; x += 32) out[o + x] = sp[x];
}
void gather_ranges(const uint8_t *d_src, const uint32_t *d_lo, const uint64_t *d_off, const uint32_t *d_n, uint32_t cap,
                   uint8_t *d_out, cudaStream_t s) {
    if (cap) NP2_K(k_gather_ranges)<<<cdiv((uint64_t)cap * 32, 128), 128, 0, s>>>(d_src, d_lo, d_off, d_n, cap, d_out);
}

/* ---- sparse host view (np2_api.cu, final phase): regions are stored in descending position (r order) */
// near[r] = 1 for every RECH region and every region whose DP index range [a, b) reaches into the window
// [a_R - W, b_R + W) of a RECH region R; the first and the last region are always selected (FASTA header span).
__global__ void k_near_mark(const uint32_t *__restrict__ cnt, const uint8_t *__restrict__ lable,
                            const uint32_t *__restrict__ ra, const uint32_t *__restrict__ rb, uint8_t *__restrict__ near) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nreg = cnt[C_NREG], N = cnt[C_N];
    if (cnt[C_ABORT] || r >= nreg) return;
    if (r == 0 || r == nreg - 1) near[r] = 1;
    if (!(lable[r] & 0x20)) return;  // LABLE_RECH
    near[r] = 1;
    const uint32_t lo = ra[r] > kRecheckWindow ? ra[r] - kRecheckWindow : 0;
    const uint64_t hi = min((uint64_t)rb[r] + kRecheckWindow, (uint64_t)N);
    for (uint32_t x = r + 1; x < nreg && rb[x] > lo; x++) near[x] = 1;  // towards lower positions
    for (uint32_t x = r; x-- > 0 && ra[x] < hi;) near[x] = 1;           // towards higher positions
}
// window of DP bases around each RECH region, in ascending position (q = nreg - 1 - r)
__global__ void k_window_sizes(const uint32_t *__restrict__ cnt, const uint8_t *__restrict__ lable,
                               const uint32_t *__restrict__ ra, const uint32_t *__restrict__ rb,
                               uint32_t *__restrict__ win_lo, uint32_t *__restrict__ win_len) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nreg = cnt[C_NREG], N = cnt[C_N];
    if (cnt[C_ABORT] || q >= nreg) return;
    const uint32_t r = nreg - 1 - q;
    uint32_t lo = 0, len = 0;
    if (lable[r] & 0x20) {
        lo = ra[r] > kRecheckWindow ? ra[r] - kRecheckWindow : 0;
        len = (uint32_t)min((uint64_t)rb[r] + kRecheckWindow, (uint64_t)N) - lo;
    }
    win_lo[q] = lo;
    win_len[q] = len;
}
__global__ void k_sub_meta(const uint32_t *__restrict__ sub, const uint32_t *__restrict__ cnt,
                           const uint32_t *__restrict__ start, const uint32_t *__restrict__ end,
                           const uint32_t *__restrict__ ra, const uint32_t *__restrict__ rb,
                           const uint8_t *__restrict__ lable, const uint32_t *__restrict__ seed_len,
                           const uint64_t *__restrict__ seed_off, const uint32_t *__restrict__ nsurv,
                           const uint32_t *__restrict__ ent_off, const uint64_t *__restrict__ q_seedoff, SubMeta o) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nreg = cnt[C_NREG];
    if (cnt[C_ABORT] || i >= cnt[C_NSUB]) return;
    const uint32_t r = sub[i];
    o.start[i] = start[r];
    o.end[i] = end[r];
    o.a[i] = ra[r];
    o.b[i] = rb[r];
    o.lable[i] = lable[r];
    o.seed_len[i] = seed_len[r];
    o.seed_off[i] = seed_off[r];
    o.nsurv[i] = nsurv[r];
    o.ent_off[i] = ent_off[r];
    o.q_seedoff[i] = q_seedoff[nreg - 1 - r];
}
__global__ void k_seed_scatter(uint32_t n, const uint32_t *__restrict__ r, const uint64_t *__restrict__ off,
                               const uint32_t *__restrict__ len, uint64_t *__restrict__ seed_off,
                               uint32_t *__restrict__ seed_len) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    seed_off[r[i]] = off[i];
    seed_len[r[i]] = len[i];
}
void near_mark(const uint32_t *d_cnt, uint32_t cap_reg, const uint8_t *d_lable, const uint32_t *d_a, const uint32_t *d_b,
               uint8_t *d_near, cudaStream_t s) {
    if (cap_reg) NP2_K(k_near_mark)<<<cdiv(cap_reg, 128), 128, 0, s>>>(d_cnt, d_lable, d_a, d_b, d_near);
}
void window_sizes(const uint32_t *d_cnt, uint32_t cap_reg, const uint8_t *d_lable, const uint32_t *d_a, const uint32_t *d_b,
                  uint32_t *d_win_lo_q, uint32_t *d_win_len_q, cudaStream_t s) {
    if (cap_reg) NP2_K(k_window_sizes)<<<cdiv(cap_reg, 128), 128, 0, s>>>(d_cnt, d_lable, d_a, d_b, d_win_lo_q, d_win_len_q);
}
void near_select(const uint8_t *d_near, uint32_t cap_reg, uint32_t *d_sub, CountsDev cd, ScanPool &pool, cudaStream_t s) {
    ScanSelect<PredFlagU8> f;
    f.pred = PredFlagU8{d_near};
    f.out = d_sub;
    f.count = cd.c + C_NSUB;
    f.cap = cap_reg;
    f.abort = nullptr;
    scan_launch(f, cd.c + C_NREG, 0, cap_reg, pool, s, cd.c + C_ABORT);
}
void sub_meta_gather(const uint32_t *d_sub, const uint32_t *d_cnt, uint32_t cap_reg, const uint32_t *d_start,
                     const uint32_t *d_end, const uint32_t *d_a, const uint32_t *d_b, const uint8_t *d_lable,
                     const uint32_t *d_seed_len, const uint64_t *d_seed_off, const uint32_t *d_nsurv,
                     const uint32_t *d_ent_off, const uint64_t *d_q_seedoff, SubMeta out, cudaStream_t s) {
    if (cap_reg)
        NP2_K(k_sub_meta)<<<cdiv(cap_reg, 128), 128, 0, s>>>(d_sub, d_cnt, d_start, d_end, d_a, d_b, d_lable, d_seed_len,
                                                              d_seed_off, d_nsurv, d_ent_off, d_q_seedoff, out);
}
void seed_scatter(uint32_t n, const uint32_t *d_r, const uint64_t *d_off, const uint32_t *d_len, uint64_t *d_seed_off,
                  uint32_t *d_seed_len, cudaStream_t s) {
    if (n) NP2_K(k_seed_scatter)<<<cdiv(n, 128), 128, 0, s>>>(n, d_r, d_off, d_len, d_seed_off, d_seed_len);
}

void assemble_sizes(AssembleDev a, uint32_t cap_reg, cudaStream_t s) {
    if (cap_reg) NP2_K(k_patch_sizes)<<<cdiv(cap_reg, 128), 128, 0, s>>>(a);
}
void region_scan_u32(const uint32_t *d_in, uint32_t *d_out, uint32_t cap_reg, int c_slot, CountsDev cd, ScanPool &pool,
                     cudaStream_t s) {
    ScanOffsets<uint32_t, uint32_t> f;
    f.in = d_in;
    f.out = d_out;
    f.c_slot = c_slot >= 0 ? cd.c + c_slot : nullptr;
    f.q_slot = nullptr;
    f.cap = ~0ULL;
    f.abort = nullptr;
    scan_launch(f, cd.c + C_NREG, 0, cap_reg, pool, s, cd.c + C_ABORT);
}
void region_scan_u64(const uint32_t *d_in, uint64_t *d_out, uint32_t cap_reg, int q_slot, CountsDev cd, ScanPool &pool,
                     cudaStream_t s) {
    ScanOffsets<uint32_t, uint64_t> f;
    f.in = d_in;
    f.out = d_out;
    f.c_slot = nullptr;
    f.q_slot = q_slot >= 0 ? cd.q + q_slot : nullptr;
    f.cap = ~0ULL;
    f.abort = nullptr;
    scan_launch(f, cd.c + C_NREG, 0, cap_reg, pool, s, cd.c + C_ABORT);
}
void region_scan_i64(const long long *d_in, long long *d_out, uint32_t cap_reg, int q_slot, CountsDev cd, ScanPool &pool,
                     cudaStream_t s) {
    ScanOffsetsI64 f;
    f.in = d_in;
    f.out = d_out;
    f.q_slot = q_slot >= 0 ? cd.q + q_slot : nullptr;
    scan_launch(f, cd.c + C_NREG, 0, cap_reg, pool, s, cd.c + C_ABORT);
}
void assemble_seed_gather(AssembleDev a, uint8_t *d_out, uint32_t cap_reg, cudaStream_t s) {
    if (cap_reg) NP2_K(k_seed_gather)<<<cdiv((uint64_t)cap_reg * 32, 128), 128, 0, s>>>(a, d_out);
}
void rech_sizes(GenoDev g, uint32_t *d_bytes, uint32_t cap_reg, cudaStream_t s) {
    if (cap_reg) NP2_K(k_rech_sizes)<<<cdiv(cap_reg, 128), 128, 0, s>>>(g, d_bytes);
}
void rech_gather(GenoDev g, const uint32_t *d_ent_off, const uint64_t *d_byte_off, uint32_t *d_order, uint32_t *d_len,
                 uint64_t *d_pool_off, uint8_t *d_out, uint32_t cap_reg, cudaStream_t s) {
    if (cap_reg) NP2_K(k_rech_gather)<<<cdiv(cap_reg, 128), 128, 0, s>>>(g, d_ent_off, d_byte_off, d_order, d_len, d_pool_off, d_out);
}
void assemble_final(AssembleDev a, uint8_t *d_out, uint32_t cap_reg, cudaStream_t s) {
    NP2_K(k_assemble)<<<cdiv((uint64_t)(cap_reg + 1) * 32, 128), 128, 0, s>>>(a, d_out);
}

}  // namespace np2
