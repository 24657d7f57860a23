// np2_host.h — host phases of the polish path that stay on the CPU by design (small, sequential, irregular):
// record parsing/filtering, the Louvain clustering of the read agreement graph, and the k-mer re-check of the few
// regions with more than one supported allele.  Everything works on flat arrays produced/consumed by the kernels.
#pragma once
#include <stdint.h>
#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
#endif

#include <functional>
#include <memory>
#include <string>
#include <algorithm>
#include <vector>

#include "../../include/np2gpu.h"

namespace np2 {

/* ---------------------------------------------------------------- ingest */
// Host arrays that are DMA sources live in memory from these hooks (malloc/free by default; the library points them
// at page-locked allocations so that uploads are asynchronous and run at PCIe speed).
extern void *(*host_alloc_hook)(size_t);
extern void (*host_free_hook)(void *);
// One column-consuming CIGAR op as the kernels read it (one 16-byte load): first alignment column, first query base,
// reference offset from the record's pos, and the raw BAM op word (len << 4 | op).
struct alignas(16) Op {
    uint32_t col, q, t, cig;
};
#if defined(__x86_64__) || defined(_M_X64)
inline void stream_store(Op *p, const Op &v) {
    _mm_stream_si128(reinterpret_cast<__m128i *>(p), _mm_set_epi32((int)v.cig, (int)v.t, (int)v.q, (int)v.col));
}
inline void store_fence() { _mm_sfence(); }
#else
inline void stream_store(Op *p, const Op &v) { *p = v; }
inline void store_fence() {}
#endif
// Host threads one call may use for record parsing / SEQ compaction: np2_set_host_threads(), else the environment
// variable NP2_HOST_THREADS, else min(16, hardware threads).  Callers that run several contexts or ranks on one box
// divide the cores between them.
unsigned host_threads();
void set_host_threads(unsigned n);
// Runs f(0) .. f(n - 1) on a process-wide pool of persistent worker threads (the caller takes part); returns when all
// are done.  Spawning 16 std::threads per contig cost about a third of the record parse.
void parallel_for(unsigned n, const std::function<void(unsigned)> &f);
// memcpy whose destination bypasses the cache (16-byte streaming stores once dst is aligned)
void copy_streaming(void *dst, const void *src, size_t n);
template <class T>
struct HVec {  // minimal growable array of trivially copyable T on the hook allocator; capacity survives clear()
    T *p = nullptr;
    size_t n = 0, cap = 0;
    HVec() {}
    HVec(const HVec &) = delete;
    HVec &operator=(const HVec &) = delete;
    HVec(HVec &&o) noexcept : p(o.p), n(o.n), cap(o.cap) { o.p = nullptr, o.n = o.cap = 0; }
    ~HVec() {
        if (p) host_free_hook(p);
    }
    void grow(size_t need);
    // Streaming (non-temporal) store: these arrays are written once by the parse threads and then read by the GPU's
    // copy engine only.  Written through the cache, the DMA has to snoop freshly dirtied lines out of the cores'
    // caches and drops from ~30 to ~8 GB/s on the test box (profiles/microbench/h2d_small_copies.cu); call
    // np2::store_fence() before handing the buffer to CUDA.
    void push_back(T v) {
        if (n == cap) grow(n + 1);
        stream_store(&p[n++], v);
    }
    void clear() { n = 0; }
    void reserve(size_t want) {  // only when empty: growing would have to copy
        if (n == 0 && want > cap) grow(want);
    }
    size_t size() const { return n; }
};

struct Ingest {
    // candidate reads = records that pass the record-level filter (main.rs:1758-1771)
    std::vector<int32_t> all_tid, all_pos;  // every record, for the sortedness assertion (main.rs:1753-1756)
    std::vector<int32_t> rec_idx;
    std::vector<uint32_t> pos, ncols, rlen, rspan;
    std::vector<uint8_t> is_clip;
    std::vector<uint64_t> seq_off;      // byte offset of SEQ in the caller's record buffer
    std::vector<uint32_t> seq_bytes;    // (l_seq + 1) / 2
    std::vector<uint32_t> n_cig;        // raw CIGAR words of the record: they sit right in front of SEQ
    std::vector<uint32_t> op_off;       // n + 1
    std::vector<uint64_t> nib_off;      // n + 1, bytes, 16-B aligned
    std::vector<uint32_t> ck_off;       // n + 1, 32-column blocks
    uint64_t total_cols = 0;
    uint64_t n_ops = 0;
    uint32_t n_fallback = 0;            // segments that had to be re-walked sequentially (speculation miss)
    // The 16-byte op records the kernels read are expanded ON THE DEVICE from the raw CIGAR words, which travel with
    // SEQ in one span per read (np2_kernels.cu k_cigar_ops); the host only sums the CIGAR (filter, sizes).  With
    // host_ops (np2_debug_parse) the host builds the same records itself: column-consuming CIGAR ops (M,=,X,I,D) in
    // the per-segment arrays they were parsed into; back to back in file order they are the array op_off[] indexes.
    bool host_ops = false;
    struct OpChunk {
        const Op *ops;
        size_t n;
    };
    std::vector<OpChunk> op_chunks;

    struct RecOut {
        uint8_t kept = 0, is_clip = 0;
        uint32_t ncols = 0, rlen = 0, rspan = 0, n_ops = 0, seq_bytes = 0, n_cig = 0;
        uint64_t seq_off = 0;
    };
    // one byte range of the record buffer, walked and parsed by one host thread
    struct Segment {
        uint64_t start = 0, end = 0;
        bool found = false;
        std::vector<int32_t> tid, pos;
        std::vector<RecOut> ro;
        HVec<Op> ops;  // page-locked; uploaded from where it was parsed
        int64_t err_rec = -1;  // local index of the first record that fails (the walk stops there)
        const char *err_msg = nullptr;
        void reset() {
            start = end = 0;
            found = false;
            tid.clear();
            pos.clear();
            ro.clear();
            ops.clear();
            err_rec = -1;
            err_msg = nullptr;
        }
    };
    std::vector<std::unique_ptr<Segment>> segs;  // pooled between jobs (capacity reuse: fresh pages are expensive)
    void clear();
};
// throws np2::Error(NP2_ERR_FORMAT) where the reference panics (the first failing record in file order decides)
// threads = 0: one range per host thread (a single range below 8 MB); otherwise exactly that many ranges
// host_ops: also build the op records on the host (debug seam); the job path leaves that to the device
void parse_records(const uint8_t *bam, uint64_t bam_len, uint32_t tlen, const np2_opts &opt, Ingest &out,
                   unsigned threads = 0, bool host_ops = false);

// The same from the HEADS of records whose boundaries are known (np2_job_create_bgzf: the records were inflated on the
// device and stay there).  heads + head_off[i] = record i's block_size, 32 fixed bytes, read name and CIGAR words;
// rec_off[i] = the record's offset in the device's record region of region_len bytes (seq_off counts from there).
void parse_heads(const uint8_t *heads, const uint64_t *head_off, const uint64_t *rec_off, uint64_t n_rec, uint64_t region_len,
                 uint32_t tlen, const np2_opts &opt, Ingest &out, unsigned threads = 0);

/* ---------------------------------------------------------------- regions */
struct Regions {
    std::vector<uint32_t> start, end;  // reference order: descending position
    std::vector<uint32_t> a, b;        // DP-consensus index range [a, b) of the bases with start <= pos <= end
};
/* ---------------------------------------------------------------- phasing */
// phase_reads_by_lqseqs (main.rs:994-1015) + louvain.rs on the reduced agreement edges produced on the device:
// key = (min order << 32 | max order), val = sum of +-1 in the low 32 bits (signed) + number of disagreements << 32.
// Order 0 is the ref read (its pairs only feed ref_data / invalid_ids).  Returns the alignseq indices to blank.
std::vector<uint32_t> phase_reads(const uint64_t *keys, const long long *vals, uint64_t n_edges, bool asref,
                                  bool use_all_reads);
std::vector<uint32_t> phase_reads_general(const uint64_t *keys, const long long *vals, uint64_t n_edges, bool asref,
                                  bool use_all_reads);
// the same from a level-0 graph already in CSR form (built on the device); `general` is called when the flat path
// has to hand over (it must produce the answer from the original edges)
std::vector<uint32_t> phase_reads_csr(uint32_t n, const uint32_t *aoff, const uint32_t *ato, const float *aw,
                                      const uint8_t *has, const uint8_t *bad_v, const uint8_t *in_ref, const float *ref_w,
                                      bool asref, const std::function<std::vector<uint32_t>()> &general);
// which implementation served the calling thread's last phase_reads: 1 = flat arrays, 2 = general path
int phase_last_path();
// host milliseconds of the last flat-array call: adjacency build, vertex moves, aggregation, communities
const float *phase_last_ms();

/* ---------------------------------------------------------------- consensus patching */
const uint8_t LABLE_TEMP = 0x01, LABLE_SUCC = 0x80, LABLE_HETE = 0x40, LABLE_RECH = 0x20;  // main.rs:655-658

struct Allele {
    const uint8_t *s = nullptr;
    uint32_t len = 0;
    uint32_t order = 0;
    uint16_t kscore = 0;
    uint64_t dev_off = 0;  // where the same string lives in the device candidate pool
};
// The consensus is never rebuilt base by base: it is the DP consensus (cbase, N bases) with the index range
// [a, b) of every region replaced by that region's current sudoseed (update_consensus_with_lqseqs
// main.rs:1027-1058 replaces exactly that range).  Regions here are in ASCENDING position.
struct Patched {
    const uint8_t *cbase = nullptr;
    const uint32_t *cpos = nullptr;
    uint64_t N = 0;
    std::vector<uint32_t> start, end, a, b;
    std::vector<uint8_t> lable;
    std::vector<Allele> seed;                 // current sudoseed per region
    std::vector<std::vector<Allele>> cand;    // surviving candidates of RECH regions (retain_sort_seqs order)
    // empties the view for n regions but keeps every allocation (fresh multi-MB vectors cost page faults per contig)
    void reset(size_t n) {
        cbase = nullptr;
        cpos = nullptr;
        N = 0;
        start.resize(n);
        end.resize(n);
        a.resize(n);
        b.resize(n);
        lable.resize(n);
        seed.resize(n);
        const size_t keep = std::min(n, cand.size());
        for (size_t i = 0; i < keep; i++) cand[i].clear();
        cand.resize(n);
    }
};
// reupdate_consensus_with_lqseqs (main.rs:1060-1420), split around the device scoring call
struct Reupdate {
    struct Group {
        uint32_t sj, ej;        // range in rech
        uint64_t first_string;  // index of the group's first string
    };
    std::vector<uint32_t> rech;  // RECH regions, ascending position
    std::vector<Group> groups;
    std::vector<uint8_t> pool;   // all strings to score, concatenated
    std::vector<uint64_t> off;   // n_strings + 1
};
void reupdate_build(const Patched &pc, uint32_t k, Reupdate &ru);
void reupdate_apply(Patched &pc, const Reupdate &ru, const uint16_t *kscores, uint32_t iter_count);
// ConsensusBase.pos of the final consensus (the bases are assembled on the device): DP positions outside the
// regions, region.start for every base of a patched region (main.rs:1039-1045)
void positions(const Patched &pc, std::vector<uint32_t> &pos);

// -S / --use_secondary (np2_secondary.cpp; secondary.rs:8-158)
struct SecMap;
void secmap_scan_ids(SecMap &m, const uint8_t *bam, uint64_t len);
void secmap_scan_seqs(SecMap &m, const uint8_t *bam, uint64_t len);
uint64_t secmap_fill(const SecMap &m, const uint8_t *bam, uint64_t len, uint8_t *out, uint64_t cap);
SecMap *secmap_new();
void secmap_delete(SecMap *m);
uint64_t secmap_counts(const SecMap &m, uint64_t *n_seqs);

}  // namespace np2
