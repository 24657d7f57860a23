/*
 * np2gpu.h — C ABI of libnp2gpu.so, the B200-native (sm_100a) implementation of the
 * NextPolish2 per-contig polish path.
 *
 * The reference (Nextomics/NextPolish2 @ 283dc5a) has no plugin / FFI interface
 * (SURVEY.md §8b); the seams this library replaces are function-level:
 *
 *   np2_yak_load / np2_yak_from_arrays   <- KmerInfo::new            src/utils/kmer.rs:72-100
 *                                           (+ one-time staging of the whole dump into HBM,
 *                                            file format yak/htab.c:190-211)
 *   np2_yak_lookup[_device]              <- KmerInfo::{insert,retrieve_kmers,get}
 *                                                                      src/utils/kmer.rs:113-170
 *   np2_seq_kscore                       <- iter2kmer + to_hash + get + min
 *                                                                      src/utils/kmer.rs:255-314,
 *                                                                      src/main.rs:761-769,1300-1315
 *   np2_polish_contig                    <- the worker closure        src/main.rs:1726-1838
 *   np2_job_* (staged form + stage dumps) <- the same closure, split at the points the
 *                                            parity tests and bench.py need
 *   np2_format_fasta                     <- display_consensusbase_vec src/main.rs:607-645
 *
 * Conventions: plain pointers and sizes, no C++/torch types.  The caller owns every
 * host buffer it passes in; the library owns device memory and the buffers returned by
 * np2_job_get_* (valid until np2_job_destroy).  All functions return 0 on success or a
 * negative NP2_ERR_* code; np2_last_error() returns the message of the calling thread's
 * last failure.  The library never aborts: conditions on which the reference panics
 * (unsorted BAM, unknown CIGAR op, bad yak magic ...) are reported as errors and the CLI
 * maps them to the reference's exit behaviour.
 *
 * Threading: one context = one GPU + one stream; calls on one context must be serialised
 * by the caller, contexts are independent.  Tables are read-only once loaded and may be
 * used by jobs of ANY context on the same GPU, so a caller that wants two contigs in
 * flight per GPU (PCIe upload + record parsing of one overlapping the kernels of the
 * other) creates two contexts, drives each from its own thread and shares one set of
 * tables (this is what the np2 CLI and bench.py's e2e arm do).
 */
#ifndef NP2GPU_H
#define NP2GPU_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define NP2_OK 0
#define NP2_ERR_CUDA (-1)        /* CUDA runtime failure / no device */
#define NP2_ERR_ARG (-2)         /* bad argument */
#define NP2_ERR_IO (-3)          /* file missing / unreadable */
#define NP2_ERR_FORMAT (-4)      /* reference would panic while parsing (bad magic, BAM parse, unknown CIGAR, unsorted input) */
#define NP2_ERR_UNSUPPORTED (-5) /* valid for the reference but outside this library's scope (k>=32 as smallest table ...) */
#define NP2_ERR_INTERNAL (-6)

/* CLI options that reach the hot path (src/utils/option.rs:15-41, defaults 267-292). */
typedef struct np2_opts {
    uint32_t min_kmer_count;    /* -k 5 */
    uint32_t iter_count;        /* -i 2 */
    uint32_t model;             /* -m: 0 = "ref", 1 = "len" */
    uint32_t min_read_len;      /* -l 1000 */
    uint64_t min_ctg_len;       /* -L 1000000 */
    int32_t  max_indel_len;     /* -n 20 */
    uint32_t use_supplementary; /* -s */
    uint32_t use_secondary;     /* -S: records must have been through np2_secmap_fill */
    uint32_t use_all_reads;     /* -r */
    uint32_t min_map_len;       /* integer part of -a 500.5 */
    float    min_map_fra;       /* fractional part of -a 500.5 */
    int32_t  min_map_qual;      /* -q 1 */
    uint32_t max_clip_len;      /* -c 100 */
    uint32_t uppercase;         /* -u (np2_format_fasta only) */
    uint32_t out_pos;           /* --out_pos (np2_format_fasta only) */
    uint32_t reserved;
} np2_opts;

typedef struct np2_ctx np2_ctx;
typedef struct np2_table np2_table;
typedef struct np2_job np2_job;

const char *np2_last_error(void);
void np2_opts_default(np2_opts *o); /* option.rs:267-292 */

/* host threads one call may use for record parsing and SEQ compaction (0 = NP2_HOST_THREADS from the environment, else
 * min(16, hardware threads)); process-wide.  A caller running several contexts or ranks on one box divides the cores. */
void np2_set_host_threads(uint32_t n);
/* per-stage CUDA-event timers behind np2_job_get_timings (default on, NP2_STAGE_TIMING=0 in the environment turns them
 * off): off keeps only "total" and the host phases, and saves two submissions per stage; process-wide. */
void np2_set_stage_timing(int on);

int np2_device_count(void); /* visible CUDA devices (0 when there is none / no driver) */
int np2_ctx_create(int device, np2_ctx **out);
void np2_ctx_destroy(np2_ctx *ctx);

/* ---- yak tables, staged once into HBM ---- */
int np2_yak_load(np2_ctx *ctx, const char *path, np2_table **out);
/* hashes: full 64-bit yak hashes (low 10 bits = sub-table id), counts: 10-bit counts */
int np2_yak_from_arrays(np2_ctx *ctx, uint32_t k, const uint64_t *hashes, const uint16_t *counts, uint64_t n,
                        np2_table **out);
void np2_yak_free(np2_table *t);
uint32_t np2_yak_k(const np2_table *t);
uint64_t np2_yak_size(const np2_table *t);
uint64_t np2_yak_device_bytes(const np2_table *t);
/* Replicating a staged table to the other GPUs of the box over NVLink instead of staging it N times over PCIe
 * (SURVEY.md 8e).  Every GPU holds a full replica (the reference clones its KmerInfo per worker thread, main.rs:1724).
 *  - np2_yak_clone: same process, another GPU: one peer copy of the device image (cudaMemcpyPeerAsync).
 *  - np2_yak_image + np2_yak_adopt: one process per GPU: the owner exposes the device image (a flat byte array), the
 *    caller moves it with its own collective (ncclBroadcast / torch.distributed.broadcast into a device buffer) and
 *    every other rank adopts the received bytes (copied; the buffer may be freed afterwards). */
int np2_yak_clone(np2_ctx *dst_ctx, const np2_table *src, np2_table **out);
int np2_yak_image(const np2_table *t, const void **d_image, uint64_t *bytes, uint32_t *buckets_per_subtable);
int np2_yak_adopt(np2_ctx *ctx, uint32_t k, uint64_t n_keys, uint32_t buckets_per_subtable, const void *d_image,
                  uint64_t bytes, np2_table **out);
/* counts[i] = stored count of hashes[i] if present and >= min_count, else 0.  Host buffers. */
int np2_yak_lookup(np2_ctx *ctx, const np2_table *t, const uint64_t *hashes, uint64_t n, uint32_t min_count,
                   uint16_t *counts);
/* Same with DEVICE buffers (inputs resident in HBM); *ms (optional) = kernel time from CUDA events on the
 * library's stream, averaged over `repeat` back-to-back launches. */
int np2_yak_lookup_device(np2_ctx *ctx, const np2_table *t, const uint64_t *d_hashes, uint64_t n, uint32_t min_count,
                          uint16_t *d_counts, uint32_t repeat, float *ms);
/* kscore of each byte string: min over its canonical k-mers of the filtered count, 0 if it has none
 * (retrieve_kmer_count main.rs:761-769; reupdate main.rs:1300-1315).  seq_off has n+1 entries.  Host buffers. */
int np2_seq_kscore(np2_ctx *ctx, const np2_table *t, const uint8_t *seqs, const uint64_t *seq_off, uint64_t n,
                   uint32_t min_count, uint16_t *kscore);

/* ---- yak count on the device: the producer of the tables (yak/count.c:28-165, htab.c:51-78, main.c:24-83) ----
 * np2_count_add takes reads as concatenated bases + n_seqs + 1 offsets (host memory), any number of times.  The counter
 * then holds, for every canonical k-mer hash (yak_hash64 for k < 32, yak_hash_long for 32 <= k < 64), its number of
 * occurrences clamped at 1023 — what `yak count` leaves in its hash tables.  np2_count_finish keeps the hashes with
 * count >= min_count (1 = plain `yak count`; 2 = `yak count -b N in.fq in.fq`, whose second pass + yak_ch_shrink(2)
 * drop the singletons the Bloom filter let through), optionally writes them in yak's dump format (readable by the
 * reference and by yak itself; the order of keys inside a sub-table is not yak's hash-table order) and / or stages
 * them as a table.  Either of dump_path / out_table may be NULL. */
typedef struct np2_counter np2_counter;
int np2_count_create(np2_ctx *ctx, uint32_t k, np2_counter **out);
int np2_count_add(np2_counter *c, const uint8_t *seqs, const uint64_t *seq_off, uint64_t n_seqs);
uint64_t np2_count_distinct(const np2_counter *c, uint64_t *n_kmers); /* distinct hashes so far; *n_kmers = k-mers seen */
int np2_count_finish(np2_counter *c, uint32_t min_count, const char *dump_path, np2_table **out_table);
void np2_count_destroy(np2_counter *c);

/* measurement aid (bench.py): mean time of n_loads independent uniformly random 32-byte sector reads over a
 * scratch buffer of buf_bytes — the measured random-read peak K5 is compared with. */
int np2_bench_gather32(np2_ctx *ctx, uint64_t buf_bytes, uint64_t n_loads, uint32_t repeat, float *ms);
/* the same with block_bytes = 32, 64 or 128: every load reads ALL sectors of a random aligned block of that size (what
 * DRAM delivers when the L2 fetches more than the one sector a probe needs) */
int np2_bench_gather(np2_ctx *ctx, uint64_t buf_bytes, uint64_t n_loads, uint32_t block_bytes, uint32_t repeat, float *ms);
/* cudaLimitMaxL2FetchGranularity of the context's device (a hint: 32, 64 or 128 bytes fetched from DRAM per L2 miss).
 * bytes = 0 only queries.  Device-wide, so a caller that changes it for a lookup-heavy phase restores *previous. */
int np2_l2_fetch_granularity(np2_ctx *ctx, uint32_t bytes, uint32_t *previous);

/* ---- -S / --use_secondary: SEQ of secondary alignments (src/utils/secondary.rs:8-158, main.rs:1775-1788) ----
 * Secondary records carry no SEQ.  Like the reference, the caller makes two passes over EVERY contig's records before
 * polishing: np2_secmap_scan_ids (names of secondary records; retrieve_secondary_ids secondary.rs:8-66), then
 * np2_secmap_scan_seqs (SEQ of the primary record of each such name, in the read's original orientation;
 * retrieve_secondary_seq_from_bam secondary.rs:85-150; two primaries with one name -> NP2_ERR_FORMAT like the
 * reference's assert).  np2_secmap_fill then rewrites one contig's record blob so that every secondary record holds
 * its SEQ (reverse-complemented when the record is on the reverse strand, A<->T C<->G only, main.rs:1776-1783) with
 * 0xFF QUAL; the result is what np2_polish_contig takes when opts.use_secondary is set.  A secondary record whose
 * name has no primary keeps an empty SEQ: if it passes the filter the polish fails (the reference panics there).
 * np2_secmap_fill: out may be NULL / cap 0 to size the buffer; *need = bytes required.  Host only, no device. */
typedef struct np2_secmap np2_secmap;
int np2_secmap_create(np2_secmap **out);
void np2_secmap_destroy(np2_secmap *m);
int np2_secmap_scan_ids(np2_secmap *m, const uint8_t *bam, uint64_t bam_len);
int np2_secmap_scan_seqs(np2_secmap *m, const uint8_t *bam, uint64_t bam_len);
int np2_secmap_fill(const np2_secmap *m, const uint8_t *bam, uint64_t bam_len, uint8_t *out, uint64_t cap, uint64_t *need);
uint64_t np2_secmap_size(const np2_secmap *m, uint64_t *n_seqs); /* #secondary names; *n_seqs = #recovered SEQs */

/* ---- page-locked host buffers ----
 * Record buffers handed to np2_polish_contig / np2_job_create may live in any host memory.  When they are
 * page-locked (from np2_host_alloc, cudaHostAlloc, cudaHostRegister, torch pin_memory ...) the device gathers the
 * CIGAR words + SEQ field of every kept record straight out of them over PCIe and QUAL / names / tags never cross the
 * bus; pageable buffers are compacted by host threads into a pinned ring first.  Same results either way. */
int np2_host_alloc(uint64_t bytes, void **out);
void np2_host_free(void *p);

/* ---- BGZF inflate on the device: the decode layer under the reference's bam::IndexedReader::fetch + records()
 * (src/main.rs:1745-1757; rust-htslib -> htslib bgzf_read_block -> zlib inflate of every <= 64 KiB member) ----
 * comp/comp_len: host buffer holding the compressed members (e.g. the memory-mapped BAM file; pageable or page-locked).
 * Member i's raw DEFLATE payload is comp[payload_off[i] .. + payload_len[i]) (the bytes between the BGZF header and the
 * CRC32/ISIZE trailer) and inflates to isize[i] bytes (the trailer's ISIZE, <= 65536).  The members are inflated back to
 * back in the order given, one warp per member, and bytes [skip, skip + out_len) of that concatenation are written to
 * `out` (host; page-locked memory from np2_host_alloc makes the copy a DMA) — a contig's records usually start and end
 * inside a member.  Only the bytes between the first and the last payload cross the link, compressed.  CRC32 is not
 * checked (neither does the inflate call of the host path); a member that is no valid DEFLATE stream of exactly its
 * ISIZE gives NP2_ERR_FORMAT.  *kernel_ms (optional): device time of the inflate kernel from CUDA events. */
int np2_bgzf_inflate(np2_ctx *ctx, const uint8_t *comp, uint64_t comp_len, const uint64_t *payload_off,
                     const uint32_t *payload_len, const uint32_t *isize, uint32_t n_members, uint64_t skip, uint64_t out_len,
                     uint8_t *out, float *kernel_ms);

/* ---- per-contig polish ----
 * tseq/tlen : contig sequence (raw FASTA bytes, case preserved)
 * bam       : this contig's BAM alignment records, concatenated in file order, each with its block_size prefix
 * tables    : one or more tables (sorted by k internally, option.rs:238)
 * The consensus comes back as parallel arrays (pos, base) = Vec<ConsensusBase> (main.rs:591-596). */
int np2_polish_contig(np2_ctx *ctx, const uint8_t *tseq, uint32_t tlen, const uint8_t *bam, uint64_t bam_len,
                      np2_table *const *tables, uint32_t n_tables, const np2_opts *opts, np2_job **out);

/* staged form: create parses + filters the records on the host (main.rs:1758-1771) and ENQUEUES the upload of the
 * inputs (it returns while the transfer runs); upload waits until they are in HBM; run executes the device pipeline +
 * host phases from resident inputs.  A caller that keeps two contexts per worker thread can therefore create the job of
 * its next contig before it runs the current one: the link and the GPU are then busy at the same time. */
int np2_job_create(np2_ctx *ctx, const uint8_t *tseq, uint32_t tlen, const uint8_t *bam, uint64_t bam_len,
                   np2_table *const *tables, uint32_t n_tables, const np2_opts *opts, np2_job **out);
/* The same from the contig's BGZF members (arguments as np2_bgzf_inflate; [skip, skip + rec_len) of the inflated
 * members = the contig's records): the members are inflated on the device and the records never come to the host.  The
 * record boundaries are found on the device (guessed per 64 KiB chunk, walked, joined on the host exactly like the host
 * parser joins its byte ranges) and only what the filter reads of a record — its fixed fields and CIGAR words, ~2 % of
 * the bytes — comes down; the SEQ + CIGAR spans are gathered device to device.  Same result as np2_bgzf_inflate +
 * np2_job_create, one third of the link traffic and no page-locked record buffer.  Replaces fetch + records() + the
 * filter of the worker closure (main.rs:1745-1771).  Not for -S (np2_secmap_fill rewrites the records on the host). */
int np2_job_create_bgzf(np2_ctx *ctx, const uint8_t *tseq, uint32_t tlen, const uint8_t *comp, uint64_t comp_len,
                        const uint64_t *payload_off, const uint32_t *payload_len, const uint32_t *isize, uint32_t n_members,
                        uint64_t skip, uint64_t rec_len, np2_table *const *tables, uint32_t n_tables, const np2_opts *opts,
                        np2_job **out);
int np2_job_upload(np2_job *job);
/* tseq and bam must stay valid (and unchanged) until np2_job_upload / np2_polish_contig returns.
 * 1 = SEQ gathered by the device from page-locked records, 2 = compacted on the host, 3 = gathered device to device from
 * records inflated there (np2_job_create_bgzf), 0 = contig below min_ctg_len */
int np2_job_ingest_path(const np2_job *job);
/* dump_iter >= 0: keep that iteration's intermediates on the host for the np2_job_get_* stage getters */
int np2_job_run(np2_job *job, int32_t dump_iter);
void np2_job_destroy(np2_job *job);

/* pos may be NULL: the per-base positions (40 MB for a 10 Mbp contig) are only built when asked for;
 * np2_job_get_span gives the first/last position the FASTA header prints (main.rs:627-632). */
uint64_t np2_job_get_consensus(np2_job *job, const uint32_t **pos, const uint8_t **base);
uint64_t np2_job_get_span(np2_job *job, uint32_t *first_pos, uint32_t *last_pos);
/* stage dumps (same shapes as the oracle's getters) */
uint64_t np2_job_get_reads(np2_job *job, const int32_t **rec_idx, const uint32_t **t_s, const uint32_t **t_e,
                           const uint64_t **nib_off, const uint8_t **nib, const uint8_t **blank_after_clip);
uint64_t np2_job_get_msa(np2_job *job, const uint64_t **off, const uint16_t **bases, const uint16_t **delta,
                         const uint32_t **count, const uint32_t **besti);
uint64_t np2_job_get_dp_consensus(np2_job *job, const uint32_t **pos, const uint8_t **base, const uint8_t **flags);
uint64_t np2_job_get_regions(np2_job *job, const uint32_t **start, const uint32_t **end, const uint8_t **lable);
uint64_t np2_job_get_candidates(np2_job *job, const uint64_t **roff, const uint32_t **order, const uint16_t **kscore,
                                const uint64_t **kmer, const uint64_t **seq_off, const uint8_t **seq);
uint64_t np2_job_get_dropped(np2_job *job, const uint32_t **ids);
/* pair weights of the dumped (non-final) iteration, the output of the pair loop of phase_reads_by_lqseqs
 * (main.rs:953-992) before anything is derived from it: keys = a << 32 | b (read orders, a < b, a = 0 is the ref read),
 * ascending; vals = #heterozygous regions where the two reads agree + #where they differ * (2^32 - 1) */
uint64_t np2_job_get_pair_weights(np2_job *job, const uint64_t **keys, const int64_t **vals);

/* measurement: per-stage device time of the last np2_job_run (CUDA events on the library's stream), kernel
 * launch count, bytes moved.  names: NUL-separated stage names; returns the number of stages. */
uint32_t np2_job_get_timings(np2_job *job, const char **names, const float **ms, const uint32_t **launches);
void np2_job_get_traffic(np2_job *job, uint64_t *h2d_bytes, uint64_t *d2h_bytes, uint64_t *n_kernel_launches,
                         uint64_t *n_alignment_columns, uint64_t *n_probes);

/* sizes of the last np2_job_run (last pass that was built): out[0] non-reference 3-mer records, out[1] distinct
 * non-reference 3-mers (Msa entries besides the reference's), out[2] runs of multi-entry positions, out[3] DP consensus
 * bases, out[4] LQ regions, out[5] (read, region) pairs, out[6] read pairs with a non-zero agreement weight,
 * out[7] passes built from scratch; since the job was created: out[8] passes that ran with speculative capacities
 * (sizes kept on the device, no read-back until the host needs data), out[9] passes that had to be repeated with
 * exact sizes; out[10] host synchronisations of the last run; out[11] reserved. */
void np2_job_get_stats(np2_job *job, uint64_t out[12]);

/* Test seam (host only, no device needed): parses + filters a record buffer (main.rs:1758-1771, 386-440) split into
 * `threads` speculative byte ranges (0 = automatic) and returns a digest of everything the parse produces:
 * out[0] records, out[1] kept reads, out[2] column-consuming CIGAR ops, out[3] alignment columns,
 * out[4] ranges re-walked sequentially because the guessed record boundary was wrong, out[5] FNV-1a of all arrays.
 * threads: low 16 bits = byte ranges (0 = default); bit 16 = parse as np2_job_create does (the CIGAR is only summed, the
 * op records are expanded on the device) and digest the per-record arrays only; bit 17 = build the op records on the
 * host (as without flags) but digest the per-record arrays only, so that the two digests can be compared; bit 18 = the
 * parse np2_job_create_bgzf runs (only the heads of the records, gathered here by a plain walk, and their offsets are
 * read), per-record arrays only: same digest as bit 16. */
int np2_debug_parse(const uint8_t *bam, uint64_t bam_len, uint32_t tlen, const np2_opts *opts, uint32_t threads,
                    uint64_t out[6]);

/* Test seam (host only): the host half of phase_reads_by_lqseqs (main.rs:994-1015) + louvain.rs:59-356 on reduced
 * agreement edges.  keys[e] = a << 32 | b (read orders, a < b, sorted ascending; a = 0 is the ref read),
 * vals[e] = sum over shared HETE regions of (+1 if the two reads agree, (1 << 32) - 1 if they differ).
 * dropped: read orders to blank, ascending (up to cap are written); *n_dropped = how many there are;
 * *path (optional) = 1 when the flat-array implementation served the call, 2 when it handed over to the general
 * one (a community with negative internal weight, louvain.rs:136-165). */
int np2_debug_phase(const uint64_t *keys, const int64_t *vals, uint64_t n_edges, uint32_t model, uint32_t use_all_reads,
                    uint32_t *dropped, uint64_t cap, uint64_t *n_dropped, uint32_t *path);

/* FASTA record exactly as display_consensusbase_vec prints it (main.rs:607-645); returns bytes needed. */
uint64_t np2_format_fasta(const char *tid, const uint32_t *pos, const uint8_t *base, uint64_t n, int uppercase,
                          int out_pos, uint8_t *out, uint64_t cap);

#ifdef __cplusplus
}
#endif
#endif
