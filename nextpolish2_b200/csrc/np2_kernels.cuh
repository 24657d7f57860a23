// np2_kernels.cuh — launch wrappers of the sm_100a kernels (definitions in np2_kernels.cu).
#pragma once
#include "np2_common.cuh"
#include "np2_scan.cuh"

namespace np2 {

/* ------------------------------------------------------------------ device-resident counts
 * Every size that one kernel produces and the next ones consume (records, groups, runs, consensus bases, events,
 * regions, pairs ...) lives in ONE small device block; kernels read their bounds from it and are launched over a
 * host-side CAPACITY.  The host therefore does not have to synchronise to size the next launch: capacities come from
 * the previous run (or are read back one by one on a first, "exact" run), the block is fetched once where the host
 * needs data anyway, and a producer that finds more than its capacity sets C_ABORT, on which every later kernel
 * returns at once (the host then repeats the pass in exact mode). */
enum Cnt : int {
    C_ABORT = 0,  // a count exceeded its capacity: everything after is skipped
    C_NREC,       // non-reference 3-mer records
    C_G,          // distinct non-reference 3-mers
    C_NRUNS,      // runs of multi-entry positions
    C_N,          // DP consensus bases
    C_NEV,        // consensus bases with qv < 95 or coverage < 2
    C_NCAND,      // closing events = candidate regions
    C_NREG,       // LQ regions
    C_NPAIRS,     // (read, region) pairs
    C_NLONG,      // candidates longer than k
    C_NU,         // read pairs with a non-zero agreement weight
    C_NSUB,       // regions the host looks at in the final phase
    C_BESTLAST,   // entry chosen at p = L - 1 when the last position is inside a run
    C_GERR,       // genotype-rule error (reference would panic)
    C_PERR,       // read pair outside its index window
    C_NENT,       // survivors of the RECH regions
    C_NLONGRUN,   // runs whose DP is done by a whole warp (k_dp_runs_long)
    C_NDIR,       // directed edges of the level-0 phasing graph (entries of its CSR adjacency)
    C_COUNT = 32
};
enum Cnt64 : int {
    Q_TOTAL = 0,  // score of the best path (main.rs:1680)
    Q_POOL,       // bytes of candidate strings
    Q_EDGES,      // candidate pairs of the heterozygous regions
    Q_SEEDS,      // bytes of seed strings the host wants
    Q_RECH,       // bytes of survivor strings
    Q_WIN,        // bytes of DP-base windows
    Q_SHIFT,      // total length change of all patches (signed)
    Q_COUNT = 16
};
struct CountsHost {
    uint32_t c[C_COUNT];
    unsigned long long q[Q_COUNT];
};
// device view: c = 32 x u32, q = 16 x u64 right behind
struct CountsDev {
    uint32_t *c = nullptr;
    unsigned long long *q = nullptr;
};
void counts_init(CountsDev cd, cudaStream_t s);  // all zero

/* ------------------------------------------------------------------ yak table (K5) */
constexpr int kBucketSlots = 4;  // 4 x u64 = one 32-byte DRAM sector per probe

struct TableDev {
    uint64_t *slots = nullptr;  // 1024 sub-tables x nb buckets x 4 slots; slot = (h >> 10) << 10 | count, 0 = empty
    uint32_t nb = 0;            // buckets per sub-table
    uint32_t k = 0;
    uint64_t n = 0;
};

void table_insert(const TableDev &t, const uint64_t *d_hashes, const uint16_t *d_counts, uint64_t n, int *d_err,
                  cudaStream_t s);
// keys in yak file layout: key = (h >> 10) << 10 | count, sub-table id given per range
void table_insert_filekeys(const TableDev &t, const uint64_t *d_keys, const uint32_t *d_sub_off /*1025*/, uint64_t n,
                           int *d_err, cudaStream_t s);
void table_probe(const TableDev &t, const uint64_t *d_hashes, uint64_t n, uint32_t min_count, uint16_t *d_out,
                 cudaStream_t s);
void seq_kscore(const TableDev &t, const uint8_t *d_seqs, const uint64_t *d_off, const uint32_t *d_sel, uint64_t n,
                uint32_t min_count, uint16_t *d_out, cudaStream_t s);

// bytes = 32 (one random sector per load), 64 or 128 (all sectors of a random aligned block of that size)
void gather32(const uint64_t *d_buf, uint64_t n_sectors, uint64_t n_loads, uint64_t seed, uint64_t *d_sink,
              cudaStream_t s, int bytes = 32);

/* ------------------------------------------------------------------ K0/K1 ingest */
void ref_codes(const uint8_t *d_ref, uint32_t L, uint8_t *d_code, uint32_t *d_refpk /* L/8 + 8 words */, int *d_bad,
               cudaStream_t s);

struct ReadsDev {
    uint32_t n_reads = 0;  // kept reads, index 0 = first BAM read (the ref read is implicit)
    // inputs (host-built)
    const uint32_t *pos = nullptr;
    const uint32_t *op_off = nullptr;   // n_reads + 1
    const uint64_t *seq_off = nullptr;  // byte offset of the 4-bit SEQ inside the blob
    const uint32_t *ncols = nullptr;    // alignment columns before trimming
    const uint64_t *nib_off = nullptr;  // byte offset into nib (16-B aligned), n_reads + 1
    const uint32_t *ck_off = nullptr;   // first 32-column block of the read, n_reads + 1
    const uint4 *ops = nullptr;  // per op: x = first column, y = first query base, z = ref offset, w = len << 4 | op
    const uint8_t *blob = nullptr;
    // outputs
    uint32_t *t_s = nullptr, *t_e = nullptr, *n = nullptr;  // trimmed start / inclusive end / column count (0 = no anchor)
    uint32_t *shift = nullptr;                               // first kept column (trim start)
    uint8_t *nib = nullptr;
    uint32_t *ck_tpos = nullptr;
    uint16_t *ck_delta = nullptr;
    uint32_t *ck_read = nullptr;
    uint16_t *blk_op = nullptr;  // per 32-column block: op (relative to op_off[read]) holding its first column; 0xFFFF = search
};
// K0: pull the SEQ fields out of a page-locked (mapped) record buffer; dst_off[r] is 16-B aligned + (source address & 15)
void gather_seq(const uint8_t *src_mapped, const uint64_t *d_src_off, const uint64_t *d_dst_off, const uint32_t *d_nbytes,
                uint8_t *d_dst, uint32_t n_reads, cudaStream_t s, bool src_on_device = false);
// raw CIGAR words (in front of every read's SEQ in the blob) -> the 16-byte op records of the column-consuming ops
void cigar_ops(const uint8_t *d_blob, const uint64_t *d_seq_off, const uint32_t *d_n_cig, const uint32_t *d_op_off,
               uint4 *d_ops, uint32_t n_reads, cudaStream_t s);
void trim_scan(const ReadsDev &r, const uint8_t *d_ref, uint32_t L, cudaStream_t s);
bool pack_columns(const ReadsDev &r, const uint8_t *d_ref, uint32_t n_blocks, const uint32_t *d_refpk, uint32_t *d_blk_odd,
                  cudaStream_t s);

/* ------------------------------------------------------------------ K2 pileup */
void cover_diff(const ReadsDev &r, const uint8_t *d_blank, int32_t *d_diff, cudaStream_t s);
void cover_scan(int32_t *d_cover, uint32_t n, ScanPool &pool, cudaStream_t s);  // in-place inclusive sum
// One CTA per stripe of contig positions: non-reference 3-mers found, bucketed, merged into Msa entries and ordered in
// shared memory; per position: entry offset / count, reference 3-mer count, articulation flag, bases a single-entry
// position emits.  C_NREC = records seen, C_G = entries written (reserved by one atomic per stripe).
uint32_t pileup_stripes(uint32_t L);
uint32_t pileup_stripe_width();
struct MsaDev {
    uint32_t L = 0;
    const uint32_t *cnt = nullptr;  // C_G = sparse groups
    uint32_t *sp_off = nullptr;  // L: first sparse entry of the position
    uint16_t *sp_cnt = nullptr;  // L: how many
    uint16_t *g_bases = nullptr, *g_delta = nullptr;
    uint32_t *g_count = nullptr, *g_first = nullptr, *g_besti = nullptr;
    int64_t *g_score = nullptr;
    int32_t *cover = nullptr;        // L (+1): reads (incl. ref) spanning p == Msa::coverage()
    uint32_t *dense_cnt = nullptr;   // L: count of the reference 3-mer (entry 0 for p >= 2)
    uint32_t *dense_besti = nullptr; // L
    int64_t *dense_score = nullptr;  // L
    uint8_t *multi = nullptr;        // L: more than one entry (or p < 2)
    const uint8_t *code = nullptr;   // ref codes
};
// d_first_ge: pileup_stripes(L) + 1 entries, the first read at or behind every stripe start (stripe_reads, once per job)
// d_blk_odd: one bit per 32-column block, set when the block holds more than reference 3-mers (block_flags, once per job;
//            (n_blocks + 255) / 256 * 8 words)
void stripe_reads(const ReadsDev &r, uint32_t L, uint32_t *d_first_ge, cudaStream_t s);
void block_flags(const ReadsDev &r, uint32_t n_blocks, const uint8_t *d_code, const uint32_t *d_refpk, uint32_t *d_blk_odd,
                 cudaStream_t s);
// per-stripe lists of the not-all-reference blocks (once per job): count, offsets (d_odd_off[stripes + 1], total also in
// d_odd_off[stripes]), fill
void stripe_odd_count(const ReadsDev &r, const uint32_t *d_blk_odd, const uint32_t *d_first_ge, uint32_t L, uint32_t max_span,
                      uint32_t *d_odd_cnt, cudaStream_t s);
void stripe_odd_offsets(const uint32_t *d_odd_cnt, uint32_t *d_odd_off, uint32_t L, uint32_t *d_total, ScanPool &pool,
                        cudaStream_t s);
void stripe_odd_fill(const ReadsDev &r, const uint32_t *d_blk_odd, const uint32_t *d_first_ge, uint32_t L, uint32_t max_span,
                     const uint32_t *d_odd_off, uint32_t *d_odd_list, uint32_t cap, cudaStream_t s);
void pileup_stripe(const ReadsDev &r, const uint8_t *d_blank, const uint8_t *d_code, const uint32_t *d_odd_off,
                   const uint32_t *d_odd_list, MsaDev m, uint32_t cap_g, CountsDev cd, uint32_t *d_n_emit, bool count_only,
                   cudaStream_t s);
void counts_reset_pileup(CountsDev cd, cudaStream_t s);

/* ------------------------------------------------------------------ K3 DP + consensus */
// first positions of the runs of multi-entry positions (C_NRUNS)
void runs_select(const uint8_t *d_multi, uint32_t L, uint32_t *d_run_start, uint32_t cap_runs, CountsDev cd, ScanPool &pool,
                 cudaStream_t s);
// d_long_runs: cap_runs entries, the runs handed over to the warp-cooperative kernel
void dp_runs(MsaDev m, const uint32_t *d_run_start, uint32_t cap_runs, CountsDev cd, uint32_t *d_long_runs, cudaStream_t s);
void emit_count_runs(MsaDev m, const uint32_t *d_run_start, uint32_t cap_runs, CountsDev cd, uint32_t *d_n_emit,
                     cudaStream_t s);
// emit_off = exclusive sum of n_emit (L + 1 entries), C_N = number of consensus bases; the bases of the single-entry
// positions are written by the same scan (d_pos / d_base / d_flags hold cap_n elements)
void emit_offsets(MsaDev m, const uint32_t *d_n_emit, uint32_t *d_emit_off, uint32_t *d_pos, uint8_t *d_base, uint8_t *d_flags,
                  uint32_t cap_n, CountsDev cd, ScanPool &pool, cudaStream_t s);
void emit_write(MsaDev m, const uint32_t *d_run_start, uint32_t cap_runs, CountsDev cd, const uint32_t *d_n_emit,
                const uint32_t *d_emit_off, uint32_t *d_pos, uint8_t *d_base, uint8_t *d_flags, cudaStream_t s);
// consensus indices with flags != 0, ascending (C_NEV)
void events_select(const uint8_t *d_cflags, uint32_t cap_n, uint32_t *d_events, uint32_t cap_ev, CountsDev cd,
                   ScanPool &pool, cudaStream_t s);

/* ------------------------------------------------------------------ per-region pipeline (np2_geno.cu) */
constexpr int kMaxCand = 60;  // LQSEQ_MAX_CAN_COUNT main.rs:30
struct GenoDev {
    const uint32_t *cnt = nullptr;  // C_NREG regions, C_NPAIRS (read, region) pairs
    const uint32_t *start = nullptr, *end = nullptr;  // regions, descending position (reference order)
    // per candidate read
    uint32_t *rd_s = nullptr, *rd_j = nullptr, *rd_np = nullptr, *rd_poff = nullptr;
    const uint32_t *rd_order = nullptr;  // alignseq index of the read
    // per (read, region) pair
    uint32_t *p_len = nullptr;
    uint64_t *p_kmer = nullptr;
    // per selected candidate: slot = region * kMaxCand + rank
    uint32_t *c_src = nullptr, *c_len = nullptr, *c_order = nullptr;
    uint64_t *c_kmer = nullptr, *c_off = nullptr;
    uint16_t *c_kscore = nullptr;
    uint8_t *c_rep = nullptr;
    uint8_t *pool = nullptr;
    uint32_t *long_list = nullptr;  // candidates longer than k (scored over all their k-mers); count = C_NLONG
    // per region
    uint32_t *r_ncand = nullptr, *r_bytes = nullptr, *r_nedge = nullptr, *r_seed_len = nullptr, *r_nsurv = nullptr;
    uint64_t *r_pool_off = nullptr, *r_edge_off = nullptr, *r_seed_off = nullptr;
    uint8_t *r_lable = nullptr, *r_surv = nullptr;
};
void geno_read_cursor(GenoDev g, const ReadsDev &R, const uint8_t *d_blank, cudaStream_t s);
void geno_cursor_min(uint32_t *d_rd_s, uint32_t n_reads, const uint32_t *d_abort, ScanPool &pool, cudaStream_t s);
void geno_read_ranges(GenoDev g, const ReadsDev &R, const uint8_t *d_blank, uint32_t k, cudaStream_t s);
void geno_pair_offsets(GenoDev g, uint32_t n_reads, uint32_t cap_pairs, CountsDev cd, ScanPool &pool, cudaStream_t s);
void geno_pair_scan(GenoDev g, const ReadsDev &R, uint32_t k, uint32_t cap_pairs, cudaStream_t s);
// d_first_ge / stripe_width: the per-stripe read index of the pileup (stripe_reads)
void geno_region_select(GenoDev g, const ReadsDev &R, const uint8_t *d_blank, const uint8_t *d_code, uint32_t L,
                        uint32_t k, uint32_t max_span, const uint32_t *d_first_ge, uint32_t stripe_width, uint32_t cap_reg,
                        cudaStream_t s);
void geno_pool_offsets(GenoDev g, uint32_t cap_reg, unsigned long long cap_pool, CountsDev cd, ScanPool &pool,
                       cudaStream_t s);
void geno_cand_write(GenoDev g, const ReadsDev &R, const uint8_t *d_code, uint32_t L, uint32_t k, uint32_t cap_reg,
                     cudaStream_t s);
void geno_cand_kscore(GenoDev g, const TableDev &t, uint32_t min_count, uint32_t cap_reg, CountsDev cd, cudaStream_t s);
void geno_region_hete(GenoDev g, uint32_t cap_reg, cudaStream_t s);
void geno_edge_offsets(GenoDev g, uint32_t cap_reg, CountsDev cd, ScanPool &pool, cudaStream_t s);
// pair-accumulator windows (d_W has na + 1 entries, the last one 0) from the alignseqs' record positions and ends
void geno_pair_windows(const uint32_t *d_as_pos, const uint32_t *d_as_te, uint32_t na, uint32_t *d_W, cudaStream_t s);
void geno_pair_window_offsets(const uint32_t *d_W, uint64_t *d_pair_off, uint32_t na, ScanPool &pool, cudaStream_t s);
void geno_edges_accum(GenoDev g, const uint64_t *d_pair_off, unsigned long long *d_acc, uint32_t cap_reg, CountsDev cd,
                      cudaStream_t s);
// non-zero slots in slot order (C_NU = how many)
void geno_edges_select(const unsigned long long *d_acc, uint32_t n_slots, uint32_t *d_sel, uint32_t cap_nu, CountsDev cd,
                       ScanPool &pool, cudaStream_t s);
void geno_edges_finish(const uint32_t *d_sel, uint32_t cap_nu, const uint64_t *d_pair_off, uint32_t n_ids,
                       const unsigned long long *d_acc, uint64_t *d_key, long long *d_val, CountsDev cd, cudaStream_t s);
struct PhaseDev {  // per read order (< n)
    uint8_t *has = nullptr, *bad_v = nullptr, *in_ref = nullptr;
    float *ref_w = nullptr;
};
// Level 0 of the phasing graph straight from the dense pair accumulator (no pair list, no sort): ref flags from the
// slots (0, b); then, a warp per read v, its neighbours in ascending order — the earlier reads u whose window reaches v
// (slot pair_off[u] + v - u - 1), then v's own window.  phase_adj_count writes the degrees (d_deg[n + 1]), has[] and
// C_NU (non-zero slots); the caller scans the degrees into d_aoff (total -> C_NDIR) and phase_adj_fill writes
// d_ato / d_aw.  d_as_pos: record position of every alignseq (ascending from index 1), max_span: longest reference span.
void phase_ref_acc(const unsigned long long *d_acc, const uint64_t *d_pair_off, uint32_t n, CountsDev cd, PhaseDev p,
                   bool asref, bool use_all, cudaStream_t s);
void phase_adj_count(const unsigned long long *d_acc, const uint64_t *d_pair_off, const uint32_t *d_as_pos, uint32_t n,
                     uint32_t max_span, CountsDev cd, PhaseDev p, bool use_all, uint32_t *d_deg, cudaStream_t s);
void phase_adj_offsets(const uint32_t *d_deg, uint32_t *d_aoff, uint32_t n, uint32_t cap_dir, CountsDev cd, ScanPool &pool,
                       cudaStream_t s);
void phase_adj_fill(const unsigned long long *d_acc, const uint64_t *d_pair_off, const uint32_t *d_as_pos, uint32_t n,
                    uint32_t max_span, CountsDev cd, PhaseDev p, bool use_all, const uint32_t *d_aoff, uint32_t *d_ato,
                    float *d_aw, uint32_t cap_dir, cudaStream_t s);

// have_rep: k_region_hete already ran on the same candidates of this pass (c_rep is valid)
void geno_region_seed(GenoDev g, int32_t max_indel_len, uint32_t cap_reg, CountsDev cd, bool have_rep, cudaStream_t s);

/* ------------------------------------------------------------------ regions + assembly (np2_regions.cu) */
struct RegionDev {
    const uint32_t *cnt = nullptr;     // C_N, C_NEV, C_NCAND, C_NREG
    const uint32_t *events = nullptr;  // consensus indices with flags != 0, ascending
    const uint8_t *cflags = nullptr, *cbase = nullptr;
    const uint32_t *cpos = nullptr;
    uint32_t *ev_close = nullptr;
    uint8_t *ev_boundary = nullptr, *ev_closes = nullptr;
    uint32_t *c_t = nullptr, *c_start = nullptr, *c_end = nullptr, *c_a = nullptr, *c_b = nullptr;
    uint32_t *c_head = nullptr, *c_hrank = nullptr;
    uint32_t *r_start = nullptr, *r_end = nullptr, *r_a = nullptr, *r_b = nullptr;
};
void regions_event_close(RegionDev d, uint32_t cap_ev, cudaStream_t s);
void regions_cand_select(RegionDev d, uint32_t cap_ev, uint32_t cap_cand, CountsDev cd, ScanPool &pool, cudaStream_t s);
void regions_make(RegionDev d, uint32_t cap_cand, cudaStream_t s);
void regions_rank(RegionDev d, uint32_t cap_cand, uint32_t cap_reg, CountsDev cd, ScanPool &pool, cudaStream_t s);
void regions_out(RegionDev d, uint32_t cap_cand, cudaStream_t s);

struct AssembleDev {
    const uint32_t *cnt = nullptr;  // C_NREG, C_N
    const uint8_t *cbase = nullptr, *pool = nullptr;
    const uint32_t *r_a = nullptr, *r_b = nullptr, *r_seed_len = nullptr;
    const uint64_t *r_seed_off = nullptr;
    long long *q_delta = nullptr, *q_shift = nullptr;  // q_shift: nreg + 1
    uint32_t *q_seedlen = nullptr;
    uint64_t *q_seedoff = nullptr;                      // nreg + 1
    const uint8_t *near = nullptr;  // optional (r order): only these regions' seed strings are gathered for the host
};
// ---- sparse host view: the host only looks at RECH regions and what lies within kRecheckWindow DP bases of them
constexpr uint32_t kRecheckWindow = 128;
struct SubMeta {  // compact copies for the selected regions, in the order of `sub` (ascending r = descending position)
    uint32_t *start = nullptr, *end = nullptr, *a = nullptr, *b = nullptr, *seed_len = nullptr, *nsurv = nullptr,
             *ent_off = nullptr;
    uint64_t *seed_off = nullptr, *q_seedoff = nullptr;
    uint8_t *lable = nullptr;
};
void near_mark(const uint32_t *d_cnt, uint32_t cap_reg, const uint8_t *d_lable, const uint32_t *d_a, const uint32_t *d_b,
               uint8_t *d_near, cudaStream_t s);
void window_sizes(const uint32_t *d_cnt, uint32_t cap_reg, const uint8_t *d_lable, const uint32_t *d_a, const uint32_t *d_b,
                  uint32_t *d_win_lo_q, uint32_t *d_win_len_q, cudaStream_t s);
void near_select(const uint8_t *d_near, uint32_t cap_reg, uint32_t *d_sub, CountsDev cd, ScanPool &pool, cudaStream_t s);  // C_NSUB
void sub_meta_gather(const uint32_t *d_sub, const uint32_t *d_cnt, uint32_t cap_reg, const uint32_t *d_start,
                     const uint32_t *d_end, const uint32_t *d_a, const uint32_t *d_b, const uint8_t *d_lable,
                     const uint32_t *d_seed_len, const uint64_t *d_seed_off, const uint32_t *d_nsurv,
                     const uint32_t *d_ent_off, const uint64_t *d_q_seedoff, SubMeta out, cudaStream_t s);
void seed_scatter(uint32_t n, const uint32_t *d_r, const uint64_t *d_off, const uint32_t *d_len, uint64_t *d_seed_off,
                  uint32_t *d_seed_len, cudaStream_t s);
void assemble_sizes(AssembleDev a, uint32_t cap_reg, cudaStream_t s);
// region-level offset scans over C_NREG elements (out has nreg + 1 entries); the total also goes to cd.q[q_slot] when
// q_slot >= 0 and to cd.c[c_slot] when c_slot >= 0
void region_scan_u32(const uint32_t *d_in, uint32_t *d_out, uint32_t cap_reg, int c_slot, CountsDev cd, ScanPool &pool,
                     cudaStream_t s);
void region_scan_u64(const uint32_t *d_in, uint64_t *d_out, uint32_t cap_reg, int q_slot, CountsDev cd, ScanPool &pool,
                     cudaStream_t s);
void region_scan_i64(const long long *d_in, long long *d_out, uint32_t cap_reg, int q_slot, CountsDev cd, ScanPool &pool,
                     cudaStream_t s);
void assemble_seed_gather(AssembleDev a, uint8_t *d_out, uint32_t cap_reg, cudaStream_t s);
void gather_ranges(const uint8_t *d_src, const uint32_t *d_lo, const uint64_t *d_off, const uint32_t *d_n, uint32_t cap,
                   uint8_t *d_out, cudaStream_t s);
void rech_sizes(GenoDev g, uint32_t *d_bytes, uint32_t cap_reg, cudaStream_t s);
void rech_gather(GenoDev g, const uint32_t *d_ent_off, const uint64_t *d_byte_off, uint32_t *d_order, uint32_t *d_len,
                 uint64_t *d_pool_off, uint8_t *d_out, uint32_t cap_reg, cudaStream_t s);
void assemble_final(AssembleDev a, uint8_t *d_out, uint32_t cap_reg, cudaStream_t s);

/* ------------------------------------------------------------------ yak count on the device (np2_count.cu) */
struct KmerCounts {  // distinct hashes, ascending, with their counts (clamped at 1023); device memory
    uint64_t *keys = nullptr;
    uint32_t *cnts = nullptr;
    uint64_t n = 0;
};
// adds the canonical k-mers of d_seq (len bytes; reads separated by any non-ACGT byte) to acc
void count_add(KmerCounts &acc, const uint8_t *d_seq, uint64_t len, uint32_t k, uint64_t *n_kmers, cudaStream_t s);
void count_free(KmerCounts &acc, cudaStream_t s);
// hashes with count >= min_count: compact device copies (caller frees with cudaFreeAsync) + size of every sub-table
void count_filter(const KmerCounts &acc, uint32_t min_count, uint64_t **d_keys, uint16_t **d_cnt, uint64_t *n,
                  uint32_t sub_size[1024], cudaStream_t s);
// the same keys as yak writes them ((hash >> 10) << 10 | count), grouped by sub-table, into host memory
void count_file_keys(const uint64_t *d_keys, const uint16_t *d_cnt, uint64_t n, uint64_t *h_out, cudaStream_t s);

/* ------------------------------------------------------------------ BGZF inflate on the device (np2_inflate.cu) */
// member i: raw DEFLATE payload d_comp[d_off[i] .. + d_clen[i]) -> d_out[d_out_off[i] .. + d_isize[i]); one warp each.
// d_comp needs 3 readable bytes in front of the first payload and 8 behind the last.  d_bad: two words, {0, 0xFFFFFFFF}
// before the launch -> {members that are no valid DEFLATE stream of their ISIZE, index of the first one}
void bgzf_inflate(const uint8_t *d_comp, const uint64_t *d_off, const uint32_t *d_clen, const uint64_t *d_out_off,
                  const uint32_t *d_isize, uint32_t n_members, uint8_t *d_out, uint32_t *d_bad, cudaStream_t s);
// Record boundaries of an inflated contig without bringing the records to the host (np2_inflate.cu): per 64 KiB chunk
// the guessed first record start (~0 = none), then per chunk the walk's end / record count / head bytes; after the host
// has joined the chunks (rec_base / head_base = exclusive prefix sums over the accepted chunks, ~0 for the others) the
// offsets of all records and their heads (block_size + 32 fixed bytes + read name + CIGAR words) in one compact buffer.
uint32_t rec_chunk_count(uint64_t n);
uint32_t rec_chunk_bytes();
void rec_chunk_starts(const uint8_t *d_rec, uint64_t n, uint64_t *d_start, cudaStream_t s);
void rec_chunk_count_walk(const uint8_t *d_rec, uint64_t n, const uint64_t *d_start, uint64_t *d_end, uint32_t *d_cnt,
                          uint64_t *d_hbytes, cudaStream_t s);
void rec_chunk_write_walk(const uint8_t *d_rec, uint64_t n, const uint64_t *d_start, const uint64_t *d_rec_base,
                          const uint64_t *d_head_base, uint64_t *d_rec_off, uint64_t *d_head_off, uint8_t *d_heads,
                          cudaStream_t s);

}  // namespace np2
