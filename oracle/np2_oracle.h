/*
 * np2_oracle.h — C ABI of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a CPU restatement of the reference
 * algorithm (Nextomics/NextPolish2 @ 283dc5a: src/main.rs, src/utils/kmer.rs,
 * src/utils/louvain.rs).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product library
 * (libnp2gpu.so) never links, loads or calls anything in oracle/.
 *
 * PARITY UNPINNED for the Rust path: the reference ships no unit tests, golden
 * outputs or KATs and cannot be built in this image (no cargo/rustc).  The
 * k-mer hash / yak-table lookup part IS pinned against the compiled reference
 * C code (oracle/_ref/yak, oracle/_ref/libyakref.so; see tests/test_oracle_yak.py).
 */
#ifndef NP2_ORACLE_H
#define NP2_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Same layout as np2_opts in include/np2gpu.h (defaults = option.rs:267-292). */
typedef struct np2o_opts {
    uint32_t min_kmer_count;    /* -k 5 */
    uint32_t iter_count;        /* -i 2 */
    uint32_t model;             /* -m: 0 = "ref", 1 = "len" */
    uint32_t min_read_len;      /* -l 1000 */
    uint64_t min_ctg_len;       /* -L 1000000 */
    int32_t  max_indel_len;     /* -n 20 */
    uint32_t use_supplementary; /* -s */
    uint32_t use_secondary;     /* -S (unsupported: error) */
    uint32_t use_all_reads;     /* -r */
    uint32_t min_map_len;       /* integer part of -a 500.5 */
    float    min_map_fra;       /* fractional part of -a 500.5 */
    int32_t  min_map_qual;      /* -q 1 */
    uint32_t max_clip_len;      /* -c 100 */
    uint32_t uppercase;         /* -u (output layer only) */
    uint32_t out_pos;           /* --out_pos (output layer only) */
    uint32_t reserved;
} np2o_opts;

typedef struct np2o_table np2o_table;
typedef struct np2o_job np2o_job;

const char *np2o_last_error(void);

/* ---- k-mer primitives (kmer.rs:223-314) ---- */
uint64_t np2o_yak_hash64(uint64_t key, uint64_t mask);
uint64_t np2o_yak_hash64_64(uint64_t key);
uint64_t np2o_yak_hash_long(const uint64_t x[4]);
/* iter2kmer + KmerInfo::to_hash: writes the hashed canonical k-mers of seq to out
 * (capacity cap); returns the number of k-mers. */
int64_t np2o_seq_hashes(const uint8_t *seq, uint64_t len, uint32_t k, uint64_t *out, uint64_t cap);

/* ---- yak tables (kmer.rs:62-221; file format yak/htab.c:190-211) ---- */
np2o_table *np2o_table_load(const char *path);
np2o_table *np2o_table_from_arrays(uint32_t k, const uint64_t *hashes, const uint16_t *counts, uint64_t n);
void np2o_table_destroy(np2o_table *t);
uint32_t np2o_table_k(const np2o_table *t);
uint64_t np2o_table_size(const np2o_table *t);
/* insert + retrieve_kmers + get: count if present and >= min_count, else 0. */
void np2o_table_lookup(const np2o_table *t, const uint64_t *hashes, uint64_t n, uint32_t min_count, uint16_t *out);
/* stream_scan != 0: answer every retrieve pass the reference's way — collect the
 * query set, then scan ALL table keys against it (kmer.rs:132-170). */
void np2o_table_set_stream_scan(np2o_table *t, int on);

/* ---- per-contig polish (main.rs:1726-1838) ----
 * bam: concatenated BAM alignment records of this contig in file order, each
 *      with its leading block_size.
 * dump_iter: which iteration's intermediates to keep for the getters (0-based).
 * Returns 0 on success, <0 on error (np2o_last_error()). */
np2o_job *np2o_job_create(const uint8_t *tseq, uint32_t tlen,
                          const uint8_t *bam, uint64_t bam_len,
                          const np2o_opts *opts,
                          np2o_table *const *tables, uint32_t n_tables);
int np2o_job_run(np2o_job *job, int32_t dump_iter);
void np2o_job_destroy(np2o_job *job);

/* getters: pointers stay valid until np2o_job_destroy */
/* reads kept after ingest (index 0 = ref read): rec index (-1 for ref), aln_t_s, aln_t_e, nibble offset/len */
uint64_t np2o_get_reads(np2o_job *, const int32_t **rec_idx, const uint32_t **t_s, const uint32_t **t_e,
                        const uint64_t **nib_off, const uint8_t **nib, const uint8_t **blank_after_clip);
/* Msa of iteration dump_iter: off[tlen+1], then per entry */
uint64_t np2o_get_msa(np2o_job *, const uint64_t **off, const uint16_t **bases, const uint16_t **delta,
                      const uint32_t **count, const uint32_t **besti);
/* consensus straight out of the DP (before any LqSeqs patching), ascending order; flags bit0 = qv<95, bit1 = cov<2 */
uint64_t np2o_get_dp_consensus(np2o_job *, const uint32_t **pos, const uint8_t **base, const uint8_t **flags);
/* LQ regions (reference order: descending position) */
uint64_t np2o_get_regions(np2o_job *, const uint32_t **start, const uint32_t **end, const uint8_t **lable);
/* candidates after retrieve_kmer_count: per region offsets roff[n_regions+1]; per candidate order, kscore, kmer, seq_off[n+1], seq bytes */
uint64_t np2o_get_candidates(np2o_job *, const uint64_t **roff, const uint32_t **order, const uint16_t **kscore,
                             const uint64_t **kmer, const uint64_t **seq_off, const uint8_t **seq);
/* reads blanked by phasing after each non-final iteration (concatenated), sorted ascending per iteration */
uint64_t np2o_get_dropped(np2o_job *, const uint32_t **ids);
/* pair weights of iteration dump_iter (main.rs:953-992): keys = a << 32 | b (read orders, a < b, a = 0 is the ref read),
 * ascending; vals = #heterozygous regions where the two agree + #where they differ * (2^32 - 1) */
uint64_t np2o_get_pair_weights(np2o_job *, const uint64_t **keys, const int64_t **vals);
/* test seam: phasing on pre-summed pair weights (see np2_oracle.cpp); returns the number of dropped reads or -1 */
int64_t np2o_debug_phase(const uint64_t *keys, const int64_t *vals, uint64_t n, uint32_t model, uint32_t use_all_reads,
                         uint32_t *out, uint64_t cap);
/* final consensus */
uint64_t np2o_get_consensus(np2o_job *, const uint32_t **pos, const uint8_t **base);
/* seconds spent in np2o_job_run */
double np2o_get_seconds(np2o_job *);
/* How often a result could have depended on FxHashMap iteration order (process-wide counters since the last reset):
 * out[0] phasing calls, out[1] communities declustered (louvain.rs:136-165), out[2] conflicting community pairs with
 * equal sort keys (louvain.rs:316-339), out[3] reads in those communities.  0 / 0 means the ascending-id order used
 * here cannot have changed anything. */
void np2o_order_exposure(uint64_t out[4], int reset);

/* FASTA record exactly as display_consensusbase_vec (main.rs:607-645) prints it; returns bytes written (or needed if cap too small) */
uint64_t np2o_format_fasta(const char *tid, const uint32_t *pos, const uint8_t *base, uint64_t n,
                           int uppercase, int out_pos, uint8_t *out, uint64_t cap);

#ifdef __cplusplus
}
#endif
#endif
