// np2_host.cpp — host phases of the polish path (see np2_host.h).  Citations are to the reference
// (Nextomics/NextPolish2 @ 283dc5a).  Nothing here calls into oracle/.
#include "np2_host.h"
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#endif

#include "np2_error.h"

#include <algorithm>
#include <mutex>
#include <memory>
#include <functional>
#include <condition_variable>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <stdexcept>
#include <thread>
#include <unordered_map>

namespace np2 {

[[noreturn]] static void herr(int code, const std::string &m) { throw Error(code, m); }

/* ================================================================= ingest */

static void *default_alloc(size_t n) { return malloc(n); }
static void default_free(void *p) { free(p); }
void *(*host_alloc_hook)(size_t) = default_alloc;
void (*host_free_hook)(void *) = default_free;
template <class T>
void HVec<T>::grow(size_t need) {
    size_t nc = std::max<size_t>(std::max(need, cap * 2), 65536);
    T *np_ = static_cast<T *>(host_alloc_hook(nc * sizeof(T)));
    if (!np_) herr(NP2_ERR_INTERNAL, "host allocation failed");
    if (n) memcpy(np_, p, n * sizeof(T));
    if (p) host_free_hook(p);
    p = np_;
    cap = nc;
}
template struct HVec<Op>;

static std::atomic<unsigned> g_host_threads{0};
void set_host_threads(unsigned n) { g_host_threads = n; }
unsigned host_threads() {
    unsigned n = g_host_threads.load();
    if (!n) {
        if (const char *e = getenv("NP2_HOST_THREADS")) n = (unsigned)atoi(e);
        if (!n) n = std::min(16u, std::thread::hardware_concurrency());
    }
    return std::max(1u, std::min(n, 64u));
}

namespace {
struct Pool {
    struct Batch {
        const std::function<void(unsigned)> *f;
        std::atomic<unsigned> next{0}, done{0};
        unsigned n = 0;
    };
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::vector<std::shared_ptr<Batch>> open;  // batches that still have unclaimed indices
    std::vector<std::thread> workers;
    bool stop = false;
    void ensure(unsigned want) {
        while (workers.size() < want) workers.emplace_back([this] { loop(); });
    }
    static bool run_one(Batch &b) {
        const unsigned i = b.next.fetch_add(1);
        if (i >= b.n) return false;
        (*b.f)(i);
        b.done.fetch_add(1);
        return true;
    }
    void loop() {
        for (;;) {
            std::shared_ptr<Batch> b;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_work.wait(lk, [&] { return stop || !open.empty(); });
                if (stop) return;
                b = open.back();
                if (b->next.load() >= b->n) {
                    open.pop_back();
                    continue;
                }
            }
            while (run_one(*b)) {
            }
            {
                std::lock_guard<std::mutex> lk(mu);
                cv_done.notify_all();
            }
        }
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv_work.notify_all();
        for (auto &t : workers) t.join();
    }
};
Pool &pool() {
    static Pool *p = new Pool();  // never destroyed: worker threads may outlive static destruction order otherwise
    return *p;
}
}  // namespace

void parallel_for(unsigned n, const std::function<void(unsigned)> &f) {
    if (n <= 1) {
        if (n) f(0);
        return;
    }
    Pool &p = pool();
    auto b = std::make_shared<Pool::Batch>();
    b->f = &f;
    b->n = n;
    {
        std::lock_guard<std::mutex> lk(p.mu);
        p.ensure(std::min(n - 1, 63u));
        p.open.push_back(b);
    }
    p.cv_work.notify_all();
    while (Pool::run_one(*b)) {
    }
    std::unique_lock<std::mutex> lk(p.mu);
    p.cv_done.wait(lk, [&] { return b->done.load() >= n; });
    for (size_t i = 0; i < p.open.size(); i++)
        if (p.open[i] == b) {
            p.open.erase(p.open.begin() + i);
            break;
        }
}

void copy_streaming(void *dst, const void *src, size_t n) {
#if defined(__x86_64__) || defined(_M_X64)
    uint8_t *d = static_cast<uint8_t *>(dst);
    const uint8_t *s = static_cast<const uint8_t *>(src);
    size_t head = (16 - ((uintptr_t)d & 15)) & 15;
    if (head > n) head = n;
    memcpy(d, s, head);
    d += head, s += head, n -= head;
    for (; n >= 64; n -= 64, d += 64, s += 64) {
        const __m128i a = _mm_loadu_si128((const __m128i *)s), b = _mm_loadu_si128((const __m128i *)(s + 16));
        const __m128i c = _mm_loadu_si128((const __m128i *)(s + 32)), e = _mm_loadu_si128((const __m128i *)(s + 48));
        _mm_stream_si128((__m128i *)d, a);
        _mm_stream_si128((__m128i *)(d + 16), b);
        _mm_stream_si128((__m128i *)(d + 32), c);
        _mm_stream_si128((__m128i *)(d + 48), e);
    }
    for (; n >= 16; n -= 16, d += 16, s += 16) _mm_stream_si128((__m128i *)d, _mm_loadu_si128((const __m128i *)s));
    memcpy(d, s, n);
#else
    memcpy(dst, src, n);
#endif
}

void Ingest::clear() {
    all_tid.clear();
    all_pos.clear();
    rec_idx.clear();
    pos.clear();
    ncols.clear();
    rlen.clear();
    rspan.clear();
    is_clip.clear();
    seq_off.clear();
    seq_bytes.clear();
    n_cig.clear();
    op_off.clear();
    nib_off.clear();
    ck_off.clear();
    op_chunks.clear();
    total_cols = 0;
    n_ops = 0;
    n_fallback = 0;
}

namespace {
typedef Ingest::RecOut RecOut;
typedef Ingest::Segment Segment;

inline int32_t rd32(const uint8_t *p) {
    int32_t v;
    memcpy(&v, p, 4);
    return v;
}
// Could a record (block_size prefix included) start at byte p?  Only used to pick speculative starting points; the
// join below accepts a segment only when the verified walk before it lands exactly on its start.
inline bool plausible(const uint8_t *bam, uint64_t bam_len, uint64_t p) {
    if (p + 36 > bam_len) return false;
    const int32_t bs = rd32(bam + p);
    if (bs < 32 || p + 4 + (uint64_t)bs > bam_len) return false;
    const uint8_t *r = bam + p + 4;
    const int32_t ref_id = rd32(r), pos = rd32(r + 4), l_seq = rd32(r + 16);
    const uint32_t l_name = r[8];
    uint16_t n_cig;
    memcpy(&n_cig, r + 12, 2);
    if (ref_id < -1 || pos < -1 || l_seq < 0 || l_name == 0) return false;
    if (32ull + l_name + 4ull * n_cig + ((uint64_t)l_seq + 1) / 2 > (uint64_t)bs) return false;
    return r[32 + l_name - 1] == 0;  // read_name is NUL terminated
}
uint64_t find_start(const uint8_t *bam, uint64_t bam_len, uint64_t lo, uint64_t hi) {
    for (uint64_t p = lo; p < hi; p++) {
        uint64_t x = p;
        int k = 0;
        for (; k < 4; k++) {
            if (x == bam_len) break;  // the chain reaches the end of the buffer
            if (!plausible(bam, bam_len, x)) {
                k = -1;
                break;
            }
            x += 4 + (uint64_t)rd32(bam + x);
        }
        if (k != -1) return p;
    }
    return UINT64_MAX;
}

// The CIGAR of one record summed without building anything (the op records are expanded on the device): everything the
// filter and the buffer sizes need.  Returns false when the record has anything the general loop of parse_one would
// flag or treat specially (an op the reference does not know, more query bases than SEQ holds, an alignment that runs
// past the contig, a soft clip that is not the first op but is followed by aligned columns ...): the caller then runs
// that loop, so the fast path never has to reproduce an error message or a corner case.
struct CigarSums {
    uint64_t rlen = 0, rspan = 0, ncols = 0, qs = 0, ts = 0;
    uint32_t n_ops = 0, aln_q_s = 0, aln_q_e = 0;
};
// membership masks over the op code (bit op set = the op takes part)
constexpr uint32_t kRlen = 0x1B3;   // M I S H = X   (seq_len_from_cigar(true))
constexpr uint32_t kRspan = 0x18D;  // M D N = X     (bam_endpos)
constexpr uint32_t kCol = 0x187;    // M I D = X     (alignment columns)
constexpr uint32_t kQs = 0x193;     // M I S = X     (query bases consumed)
constexpr uint32_t kTs = 0x185;     // M D = X       (contig bases consumed)
constexpr uint32_t kKnown = 0x1B7;  // M I D S H = X
constexpr uint32_t kMI = 0x183;     // M I = X       (ops whose query range is checked against l_seq)
#if defined(__x86_64__) && defined(__GNUC__)
// Eight ops per step.  Only groups of plain ops (M I D = X, each shorter than 2^24) are summed here; the caller walks
// every other group — the first one, any with a clip or an unknown op — with the scalar loop, which also keeps the
// order-dependent values (aln_q_s / aln_q_e, the query position after the last aligned op).  Returns how many ops
// (a multiple of 8, counted from `from`) were consumed before a group it does not take.
struct CigarAcc {
    uint64_t ncols, qs, ts;
    uint32_t n_ops;
};
__attribute__((target("avx2"))) inline void cigar_flush_avx2(__m256i &a_col, __m256i &a_q, __m256i &a_t, __m256i &a_n, CigarAcc &acc) {
    alignas(32) uint32_t v[4][8];
    _mm256_store_si256((__m256i *)v[0], a_col);
    _mm256_store_si256((__m256i *)v[1], a_q);
    _mm256_store_si256((__m256i *)v[2], a_t);
    _mm256_store_si256((__m256i *)v[3], a_n);
    for (int k = 0; k < 8; k++) {
        acc.ncols += v[0][k];
        acc.qs += v[1][k];
        acc.ts += v[2][k];
        acc.n_ops += v[3][k];
    }
    a_col = a_q = a_t = a_n = _mm256_setzero_si256();
}
__attribute__((target("avx2"))) inline uint32_t cigar_groups_avx2(const uint8_t *cg, uint32_t from, uint32_t n_cig, CigarAcc &acc) {
    const __m256i one = _mm256_set1_epi32(1), zero = _mm256_setzero_si256(), ones = _mm256_set1_epi32(-1);
    const __m256i m_plain = _mm256_set1_epi32((int)kCol), m_q = _mm256_set1_epi32((int)(kQs & kCol)), m_t = _mm256_set1_epi32((int)kTs);
    __m256i a_col = zero, a_q = zero, a_t = zero, a_n = zero;
    uint32_t i = from, steps = 0;
    for (; i + 8 <= n_cig; i += 8) {
        const __m256i c = _mm256_loadu_si256((const __m256i *)(cg + 4 * (size_t)i));
        const __m256i l = _mm256_srli_epi32(c, 4), bit = _mm256_sllv_epi32(one, _mm256_and_si256(c, _mm256_set1_epi32(15)));
        const __m256i plain = _mm256_cmpeq_epi32(_mm256_and_si256(bit, m_plain), bit);  // all ones where M I D = X
        const __m256i small = _mm256_cmpeq_epi32(_mm256_srli_epi32(l, 24), zero);
        if (_mm256_movemask_epi8(_mm256_and_si256(plain, small)) != -1) break;
        const __m256i isq = _mm256_cmpeq_epi32(_mm256_and_si256(bit, m_q), bit), ist = _mm256_cmpeq_epi32(_mm256_and_si256(bit, m_t), bit);
        a_col = _mm256_add_epi32(a_col, l);
        a_q = _mm256_add_epi32(a_q, _mm256_and_si256(l, isq));
        a_t = _mm256_add_epi32(a_t, _mm256_and_si256(l, ist));
        a_n = _mm256_sub_epi32(a_n, _mm256_xor_si256(_mm256_cmpeq_epi32(l, zero), ones));  // += (l != 0)
        if (++steps == 120) {  // 120 x 2^24 < 2^31 per lane
            cigar_flush_avx2(a_col, a_q, a_t, a_n, acc);
            steps = 0;
        }
    }
    cigar_flush_avx2(a_col, a_q, a_t, a_n, acc);
    return i - from;
}
inline bool have_avx2() {
    static const bool v = __builtin_cpu_supports("avx2");
    return v;
}
#else
struct CigarAcc {
    uint64_t ncols, qs, ts;
    uint32_t n_ops;
};
inline uint32_t cigar_groups_avx2(const uint8_t *, uint32_t, uint32_t, CigarAcc &) { return 0; }
inline bool have_avx2() { return false; }
#endif
inline bool cigar_sums(const uint8_t *cg, uint32_t n_cig, int32_t l_seq, int32_t pos, uint32_t tlen, CigarSums &o) {
    uint64_t rlen = 0, rspan = 0, ncols = 0, qs = 0, ts = 0, q_hi = 0;
    uint32_t n_ops = 0, unknown = 0, aln_q_s = 0, aln_q_e = 0;
    const bool vec = have_avx2();
    for (uint32_t i = 0; i < n_cig; i++) {
        if (vec && i && (i & 7) == 0 && i + 8 <= n_cig) {
            // runs of plain groups: M I = X add to rlen what they add to the query position, M D = X to rspan what
            // they add to the contig position
            CigarAcc acc{0, 0, 0, 0};
            const uint32_t took = cigar_groups_avx2(cg, i, n_cig, acc);
            if (took) {
                ncols += acc.ncols;
                qs += acc.qs;
                ts += acc.ts;
                rlen += acc.qs;
                rspan += acc.ts;
                n_ops += acc.n_ops;
                // q_hi = query position after the last M I = X op: deletions at the end of the run do not move it
                uint32_t j = i + took;
                while (j > i && ((uint32_t)rd32(cg + 4 * (size_t)(j - 1)) & 15) == 2) j--;
                if (j > i) q_hi = qs;
                i += took - 1;
                continue;
            }
        }
        const uint32_t c = (uint32_t)rd32(cg + 4 * i);
        const uint32_t l = c >> 4, op = c & 15;
        const uint32_t bit = 1u << op;  // op < 16
        unknown |= bit & ~kKnown;
        if (op == 4) {  // main.rs via parse_one: the first op sets aln_q_s, any later soft clip sets aln_q_e
            if (i == 0) aln_q_s = l;
            else aln_q_e = (uint32_t)qs;
        }
        rlen += (bit & kRlen) ? l : 0;
        rspan += (bit & kRspan) ? l : 0;
        ncols += (bit & kCol) ? l : 0;
        n_ops += ((bit & kCol) && l) ? 1u : 0u;
        qs += (bit & kQs) ? l : 0;
        ts += (bit & kTs) ? l : 0;
        q_hi = (bit & kMI) ? qs : q_hi;
    }
    if (unknown || q_hi > (uint64_t)(uint32_t)l_seq || pos < 0 || (uint64_t)(uint32_t)pos + ts > tlen || ncols >= (1ull << 32) ||
        qs >= (1ull << 32))
        return false;
    o.rlen = rlen;
    o.rspan = rspan;
    o.ncols = ncols;
    o.qs = qs;
    o.ts = ts;
    o.n_ops = n_ops;
    o.aln_q_s = aln_q_s;
    o.aln_q_e = aln_q_e;
    return true;
}

// One record: filter (main.rs:1758-1771) + fill_with_cigar bookkeeping (main.rs:386-440) without the strings.
// Returns an error message when the reference would panic on it.
// r: the record's bytes behind its block_size (which sits at r - 4); payload: where r lies in the caller's record
// buffer (what seq_off is counted from) — the same place for a parse of the records themselves, another one when only
// the heads of the records are at hand (parse_heads).
const char *parse_rec(const uint8_t *r, uint64_t payload, uint32_t tlen, const np2_opts &opt, Segment &sg, bool host_ops);
const char *parse_one(const uint8_t *bam, uint64_t payload, uint32_t tlen, const np2_opts &opt, Segment &sg, bool host_ops) {
    return parse_rec(bam + payload, payload, tlen, opt, sg, host_ops);
}
const char *parse_rec(const uint8_t *r, uint64_t payload, uint32_t tlen, const np2_opts &opt, Segment &sg, bool host_ops) {
    const int32_t bs = rd32(r - 4), ref_id = rd32(r), pos = rd32(r + 4), l_seq = rd32(r + 16);
    const uint32_t l_name = r[8], mapq = r[9];
    uint16_t n_cig, flag;
    memcpy(&n_cig, r + 12, 2);
    memcpy(&flag, r + 14, 2);
    sg.tid.push_back(ref_id);
    sg.pos.push_back(pos);
    sg.ro.emplace_back();
    if (l_seq < 0 || 32ull + l_name + 4ull * n_cig + ((uint64_t)l_seq + 1) / 2 > (uint64_t)bs)
        return "BAM/SAM parsing failed!";
    const uint8_t *cg = r + 32 + l_name;
    CigarSums cs;
    if (!host_ops && cigar_sums(cg, n_cig, l_seq, pos, tlen, cs)) {
        const int64_t span = ((flag & 4) || n_cig == 0 || cs.rspan == 0) ? 1 : (int64_t)cs.rspan;
        const int64_t need = std::max<int64_t>((int64_t)opt.min_map_len, (int64_t)((float)cs.rlen * opt.min_map_fra));
        const bool rejected = (flag & 0x404) || (int16_t)mapq <= (int16_t)opt.min_map_qual || cs.rlen <= opt.min_read_len ||
                              ((flag & 0x100) && !opt.use_secondary) || ((flag & 0x800) && !opt.use_supplementary) ||
                              span < need;
        if (rejected) return nullptr;
        const uint32_t aln_q_e = cs.aln_q_e ? cs.aln_q_e : (uint32_t)cs.qs;
        RecOut &o = sg.ro.back();
        o.kept = 1;
        o.is_clip = (uint32_t)(aln_q_e - cs.aln_q_s + opt.max_clip_len) < (uint32_t)cs.rlen ? 1 : 0;  // main.rs:1796
        o.ncols = (uint32_t)cs.ncols;
        o.rlen = (uint32_t)cs.rlen;
        o.rspan = (uint32_t)cs.rspan;
        o.n_ops = cs.n_ops;
        o.n_cig = n_cig;
        o.seq_off = payload + 32 + l_name + 4ull * n_cig;
        o.seq_bytes = ((uint32_t)l_seq + 1) / 2;
        return nullptr;
    }
    // One pass over the CIGAR: seq_len_from_cigar(true) and bam_endpos for the filter, and the
    // column-consuming ops for the kernels.  The ops are rolled back when the filter rejects the record; what the
    // reference would panic on only counts for records that pass it.
    const size_t ops0 = sg.ops.n;
    uint64_t rlen = 0, rspan = 0;
    uint32_t qs = 0, ts = 0, col = 0, aln_q_s = 0, aln_q_e = 0, n_ops = 0;
    bool first = true;
    const char *bad = nullptr;
    for (uint32_t i = 0; i < n_cig; i++) {
        const uint32_t c = (uint32_t)rd32(cg + 4 * i);
        const uint32_t l = c >> 4, op = c & 15;
        if (op == 0 || op == 1 || op == 4 || op == 5 || op == 7 || op == 8) rlen += l;
        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rspan += l;
        if (bad) continue;
        switch (op) {
            case 4:
                qs += l;
                if (first) aln_q_s = qs;
                else aln_q_e = qs - l;
                break;
            case 0: case 7: case 8: case 1: case 2:
                if (op != 2 && (uint64_t)qs + l > (uint64_t)l_seq) {
                    bad = "CIGAR consumes more query bases than SEQ holds";
                    break;
                }
                if (op != 1 && (uint64_t)(uint32_t)pos + ts + l > tlen) {
                    bad = "alignment runs past the end of the contig";
                    break;
                }
                if (l) {
                    if (host_ops) sg.ops.push_back(Op{col, qs, ts, c});
                    n_ops++;
                }
                col += l;
                if (op != 2) qs += l;
                if (op != 1) ts += l;
                break;
            case 5:
                break;
            default:
                bad = "Unknown cigar";
        }
        first = false;
    }
    const int64_t span = ((flag & 4) || n_cig == 0 || rspan == 0) ? 1 : (int64_t)rspan;
    const int64_t need = std::max<int64_t>((int64_t)opt.min_map_len, (int64_t)((float)rlen * opt.min_map_fra));
    const bool rejected = (flag & 0x404) || (int16_t)mapq <= (int16_t)opt.min_map_qual || rlen <= opt.min_read_len ||
                          ((flag & 0x100) && !opt.use_secondary) || ((flag & 0x800) && !opt.use_supplementary) ||
                          span < need;
    if (rejected || bad || pos < 0 || (uint64_t)pos > tlen) {
        sg.ops.n = ops0;
        if (rejected) return nullptr;
        if (pos < 0 || (uint64_t)pos > tlen) return "alignment starts outside the contig";
        return bad;
    }
    if (aln_q_e == 0) aln_q_e = qs;
    RecOut &o = sg.ro.back();
    o.kept = 1;
    o.is_clip = (uint32_t)(aln_q_e - aln_q_s + opt.max_clip_len) < (uint32_t)rlen ? 1 : 0;  // main.rs:1796
    o.ncols = col;
    o.rlen = (uint32_t)rlen;
    o.rspan = (uint32_t)rspan;
    o.n_ops = n_ops;
    o.n_cig = n_cig;
    o.seq_off = payload + 32 + l_name + 4ull * n_cig;
    o.seq_bytes = ((uint32_t)l_seq + 1) / 2;
    return nullptr;
}

// walk the block_size chain from byte `from` while the record starts before `limit`, parsing as it goes
void walk(const uint8_t *bam, uint64_t bam_len, uint64_t from, uint64_t limit, uint32_t tlen, const np2_opts &opt,
          Segment &sg, bool host_ops) {
    uint64_t p = from;
    sg.start = from;
    while (p < limit && p + 4 <= bam_len) {
        const int32_t bs = rd32(bam + p);
        if (bs < 32 || p + 4 + (uint64_t)bs > bam_len) {
            sg.err_rec = (int64_t)sg.ro.size();
            sg.err_msg = "BAM/SAM parsing failed!";
            break;
        }
        // The next record's address is known as soon as this block_size is: start its cache (and TLB) misses now — header,
        // name and the first CIGAR words, eight lines — so that they overlap the parse of this record instead of
        // stalling the next iteration (the walk is otherwise one chain of dependent misses through a buffer of
        // hundreds of MB: ~0.5 us per record, most of it waiting).
        {
            const uint64_t nx = p + 4 + (uint64_t)bs;
            if (nx + 512 <= bam_len)
                for (int k = 0; k < 8; k++) __builtin_prefetch(bam + nx + 64 * k, 0, 1);
        }
        const char *m = parse_one(bam, p + 4, tlen, opt, sg, host_ops);
        if (m) {
            sg.err_rec = (int64_t)sg.ro.size() - 1;
            sg.err_msg = m;
            break;
        }
        p += 4 + (uint64_t)bs;
    }
    sg.end = p;
}

// concatenate the per-record scalars of the accepted segments (small); the op arrays stay where they are
void finish_ingest(const std::vector<Segment *> &order, Ingest &out) {
    size_t nrec = 0, nk = 0;
    for (Segment *sg : order) {
        nrec += sg->ro.size();
        for (auto &o : sg->ro) nk += o.kept;
    }
    out.all_tid.reserve(nrec);
    out.all_pos.reserve(nrec);
    out.rec_idx.reserve(nk);
    out.pos.reserve(nk);
    out.ncols.reserve(nk);
    out.rlen.reserve(nk);
    out.rspan.reserve(nk);
    out.is_clip.reserve(nk);
    out.seq_off.reserve(nk);
    out.seq_bytes.reserve(nk);
    out.n_cig.reserve(nk);
    out.op_off.reserve(nk + 1);
    out.nib_off.reserve(nk + 1);
    out.ck_off.reserve(nk + 1);
    out.op_off.push_back(0);
    out.nib_off.push_back(0);
    out.ck_off.push_back(0);
    size_t rec = 0;
    for (Segment *sg : order) {
        out.all_tid.insert(out.all_tid.end(), sg->tid.begin(), sg->tid.end());
        out.all_pos.insert(out.all_pos.end(), sg->pos.begin(), sg->pos.end());
        for (size_t i = 0; i < sg->ro.size(); i++, rec++) {
            const RecOut &o = sg->ro[i];
            if (!o.kept) continue;
            out.rec_idx.push_back((int32_t)rec);
            out.pos.push_back((uint32_t)sg->pos[i]);
            out.ncols.push_back(o.ncols);
            out.rlen.push_back(o.rlen);
            out.rspan.push_back(o.rspan);
            out.is_clip.push_back(o.is_clip);
            out.seq_off.push_back(o.seq_off);
            out.seq_bytes.push_back(o.seq_bytes);
            out.n_cig.push_back(o.n_cig);
            out.op_off.push_back(out.op_off.back() + o.n_ops);
            out.nib_off.push_back(out.nib_off.back() + ((((uint64_t)o.ncols / 16 + 1) * 8 + 15) & ~15ull));
            out.ck_off.push_back(out.ck_off.back() + (o.ncols + 31) / 32);
            out.total_cols += o.ncols;
        }
        if (sg->ops.n) out.op_chunks.push_back({sg->ops.p, sg->ops.n});
    }
    for (Segment *sg : order)
        for (auto &o : sg->ro)
            if (o.kept) out.n_ops += o.n_ops;
    if (out.n_ops >= (1ull << 32)) herr(NP2_ERR_UNSUPPORTED, "more than 2^32 CIGAR operations in one contig");
}
}  // namespace

// The block_size chain is a linked list through a buffer of hundreds of MB: followed from the front it is one
// cache + TLB miss per record, strictly serial.  Instead every host thread takes a byte range, guesses the first
// record boundary inside it (a run of four plausible headers), and walks + parses from there.  A range is accepted
// only if the verified walk before it ends exactly on its guessed start; otherwise that range is re-walked
// sequentially.  So the result never depends on the guess, only the speed does.
void parse_records(const uint8_t *bam, uint64_t bam_len, uint32_t tlen, const np2_opts &opt, Ingest &out,
                   unsigned threads, bool host_ops) {
    out.clear();
    out.host_ops = host_ops;
    unsigned T = host_threads();
    if (bam_len < (8u << 20)) T = 1;
    if (threads) T = std::min(threads, 64u);
    auto &segs = out.segs;
    while (segs.size() < T) segs.emplace_back(new Segment());
    auto bound = [&](unsigned ti) { return ti >= T ? bam_len : bam_len / T * ti; };
    auto work = [&](unsigned ti) {
        Segment &sg = *segs[ti];
        sg.reset();
        // one page-locked allocation per segment instead of a chain of doublings (each one a cudaHostAlloc + copy +
        // cudaFreeHost): HiFi records carry about one column-consuming op per 400 bytes
        if (host_ops) sg.ops.reserve((bound(ti + 1) - bound(ti)) / 256 + 65536);
        const uint64_t lo = bound(ti), hi = bound(ti + 1);
        const uint64_t from = ti == 0 ? 0 : find_start(bam, bam_len, lo, hi);
        if (from == UINT64_MAX) return;
        sg.found = true;
        try {
            walk(bam, bam_len, from, hi, tlen, opt, sg, host_ops);
        } catch (const std::exception &) {  // allocation failure inside a worker thread
            sg.err_rec = (int64_t)sg.ro.size();
            sg.err_msg = "host allocation failed while parsing the records";
        }
        store_fence();  // the op arrays were written with streaming stores
    };
    parallel_for(T, work);
    // join: accept, or re-walk what the speculation missed
    std::vector<Segment *> order;
    uint64_t cur = 0;
    for (unsigned ti = 0; ti < T; ti++) {
        Segment *sg = segs[ti].get();
        const uint64_t hi = bound(ti + 1);
        if (!(sg->found && sg->start == cur)) {
            if (cur >= hi) continue;  // a record spans the whole range
            const size_t idx = T + out.n_fallback++;
            if (segs.size() <= idx) segs.emplace_back(new Segment());
            sg = segs[idx].get();
            sg->reset();
            sg->found = true;
            walk(bam, bam_len, cur, hi, tlen, opt, *sg, host_ops);
            store_fence();
        }
        order.push_back(sg);
        if (sg->err_msg) herr(NP2_ERR_FORMAT, sg->err_msg);
        cur = sg->end;
    }
    if (cur != bam_len) herr(NP2_ERR_FORMAT, "BAM/SAM parsing failed!");
    finish_ingest(order, out);
}

// Records whose boundaries are already known and of which only the HEADS are in host memory (the records themselves
// were inflated on the device and stay there: np2_job_create_bgzf).  heads + head_off[i] = record i's block_size, 32
// fixed bytes, read name and CIGAR words; rec_off[i] = where the record starts in the device's record region of
// region_len bytes (what seq_off is counted from).  The chain is checked here: every record must begin where the one
// before it ends and the last one must end with the region.
void parse_heads(const uint8_t *heads, const uint64_t *head_off, const uint64_t *rec_off, uint64_t n_rec, uint64_t region_len,
                 uint32_t tlen, const np2_opts &opt, Ingest &out, unsigned threads) {
    out.clear();
    out.host_ops = false;
    unsigned T = host_threads();
    if (n_rec < 4096) T = 1;
    if (threads) T = std::min(threads, 64u);
    T = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(T, n_rec ? n_rec : 1));
    auto &segs = out.segs;
    while (segs.size() < T) segs.emplace_back(new Segment());
    auto work = [&](unsigned ti) {
        Segment &sg = *segs[ti];
        sg.reset();
        sg.found = true;
        const uint64_t b = n_rec * ti / T, e = n_rec * (ti + 1) / T;
        for (uint64_t i = b; i < e; i++) {
            const uint8_t *h = heads + head_off[i];
            const int32_t bs = rd32(h);
            const uint64_t next = i + 1 < n_rec ? rec_off[i + 1] : region_len;
            if (bs < 32 || rec_off[i] + 4 + (uint64_t)bs != next) {
                sg.err_rec = (int64_t)sg.ro.size();
                sg.err_msg = "BAM/SAM parsing failed!";
                break;
            }
            if (i + 8 < e) __builtin_prefetch(heads + head_off[i + 8], 0, 1);
            const char *m = parse_rec(h + 4, rec_off[i] + 4, tlen, opt, sg, false);
            if (m) {
                sg.err_rec = (int64_t)sg.ro.size() - 1;
                sg.err_msg = m;
                break;
            }
        }
    };
    parallel_for(T, work);
    if (n_rec ? rec_off[0] != 0 : region_len != 0) herr(NP2_ERR_FORMAT, "BAM/SAM parsing failed!");
    std::vector<Segment *> order;
    for (unsigned ti = 0; ti < T; ti++) {
        Segment *sg = segs[ti].get();
        order.push_back(sg);
        if (sg->err_msg) herr(NP2_ERR_FORMAT, sg->err_msg);
    }
    finish_ingest(order, out);
}

/* ================================================================= phasing: graph + Louvain */

namespace {

// louvain.rs with flat arrays.  Vertex / community ids are alignseq indices (plus the few ids the decluster step
// invents), so everything is indexed by id; iteration is always in ascending id (SURVEY hard part 3: the reference
// iterates FxHashMaps, whose order is not reproducible here).
typedef std::vector<std::pair<uint32_t, float>> AdjList;  // sorted by neighbour id

struct Level {
    std::vector<uint32_t> ids;                     // vertices of this level (keys of `data`), ascending
    std::vector<AdjList> adj;                      // indexed by id
    std::vector<uint32_t> cid;                     // vertex -> community id (Node.id)
    std::vector<float> nweight;                    // Node.weight
    std::vector<std::vector<uint32_t>> members;    // Node.nodes (original vertices, sorted)
    std::map<uint32_t, std::set<uint32_t>> comm;   // community id -> vertices (may hold empty sets)
    void grow(uint32_t id) {
        if (id >= adj.size()) {
            adj.resize(id + 1);
            cid.resize(id + 1, 0);
            nweight.resize(id + 1, 0.f);
            members.resize(id + 1);
        }
    }
};

// louvain.rs:72-117
bool move_vertices(Level &lv) {
    bool moved_any = false;
    std::vector<std::pair<uint32_t, float>> acc;
    for (;;) {
        bool stop = true;
        for (uint32_t v : lv.ids) {
            const uint32_t cur = lv.cid[v];
            acc.clear();
            for (auto &e : lv.adj[v]) {
                const uint32_t c = lv.cid[e.first];
                bool found = false;
                for (auto &a : acc)
                    if (a.first == c) {
                        a.second += e.second;
                        found = true;
                        break;
                    }
                if (!found) acc.emplace_back(c, e.second);
            }
            if (acc.empty()) continue;
            uint32_t bid = acc[0].first;
            float bw = acc[0].second;
            for (auto &a : acc)
                if (a.second > bw || (a.second == bw && a.first < bid)) {  // max weight, ties -> smaller id
                    bid = a.first;
                    bw = a.second;
                }
            if (bw > 0.0f && bid != cur) {
                lv.cid[v] = bid;
                lv.comm[bid].insert(v);
                lv.comm[cur].erase(v);
                stop = false;
                moved_any = true;
            }
        }
        if (stop) break;
    }
    return moved_any;
}

// weight of a community: its vertices' own weights + half of every internal edge seen from both ends
float internal_weight(const Level &lv, const std::set<uint32_t> &nodes) {
    float w = 0.f;
    for (uint32_t n : nodes) {
        w += lv.nweight[n];
        for (auto &e : lv.adj[n])
            if (nodes.count(e.first)) w += e.second / 2.0f;
    }
    return w;
}

// louvain.rs:119-195
Level aggregate(Level &lv) {
    Level nx;
    std::set<uint32_t> nx_keys;  // `communities` / `node` keys of the next level
    std::vector<uint32_t> decluster;
    for (auto &kv : lv.comm) {
        if (kv.second.empty()) continue;
        const uint32_t id = kv.first;
        float w = 0.f;
        const bool fast = true;
        (void)fast;
        for (uint32_t n : kv.second) {
            w += lv.nweight[n];
            for (auto &e : lv.adj[n])
                if (lv.cid[e.first] == id && kv.second.count(e.first)) w += e.second / 2.0f;
        }
        if (w < 0.f) {
            decluster.push_back(id);
            continue;
        }
        nx.grow(id);
        nx.cid[id] = id;
        nx.nweight[id] = w;
        auto &mm = nx.members[id];
        for (uint32_t n : kv.second) mm.insert(mm.end(), lv.members[n].begin(), lv.members[n].end());
        std::sort(mm.begin(), mm.end());
        mm.erase(std::unique(mm.begin(), mm.end()), mm.end());
        nx.comm[id] = {id};
        nx_keys.insert(id);
    }
    for (uint32_t id : decluster) {  // communities with negative internal weight fall apart again
        auto it = lv.comm.find(id);
        if (it == lv.comm.end()) herr(NP2_ERR_FORMAT, "louvain: declustered community vanished (reference would panic)");
        std::set<uint32_t> nodes = std::move(it->second);
        lv.comm.erase(it);
        for (uint32_t nid : nodes) {
            uint32_t nn = nid;
            while (nx_keys.count(nn)) nn++;
            nx.grow(nn);
            nx.cid[nn] = nn;
            nx.nweight[nn] = lv.nweight[nid];
            nx.members[nn] = lv.members[nid];
            nx.comm[nn] = {nn};
            nx_keys.insert(nn);
            lv.comm[nn] = {nid};
        }
    }
    // edges between the (possibly re-keyed) communities: one pass over the old edges via vertex -> owner
    std::vector<uint32_t> owner(lv.adj.size(), 0xFFFFFFFFu);
    bool clean = true;
    for (auto &kv : lv.comm)
        for (uint32_t n : kv.second) {
            if (owner[n] != 0xFFFFFFFFu) clean = false;
            owner[n] = kv.first;
        }
    std::map<std::pair<uint32_t, uint32_t>, float> sum;
    if (clean) {
        for (auto &kv : lv.comm)
            for (uint32_t v : kv.second)
                for (auto &e : lv.adj[v]) {
                    const uint32_t o = owner[e.first];
                    if (o == 0xFFFFFFFFu || !(o > kv.first)) continue;
                    sum[{kv.first, o}] += e.second;
                }
    } else {  // a vertex listed in two communities (reference quirk): literal pairwise form
        for (auto &c1 : lv.comm) {
            if (c1.second.empty()) continue;
            for (auto &c2 : lv.comm) {
                if (!(c2.first > c1.first) || c2.second.empty()) continue;
                float w = 0.f;
                bool any = false;
                for (uint32_t v : c1.second)
                    for (auto &e : lv.adj[v])
                        if (c2.second.count(e.first)) {
                            w += e.second;
                            any = true;
                        }
                if (any) sum[{c1.first, c2.first}] = w;
            }
        }
    }
    for (uint32_t id : nx_keys) nx.grow(id);
    for (auto &kv : sum) {
        if (kv.second == 0.f) continue;
        nx.grow(std::max(kv.first.first, kv.first.second));
        nx.adj[kv.first.first].emplace_back(kv.first.second, kv.second);
        nx.adj[kv.first.second].emplace_back(kv.first.first, kv.second);
    }
    // `data` keys of the next level = communities that have at least one non-zero edge (louvain.rs:183-186)
    for (uint32_t id = 0; id < nx.adj.size(); id++)
        if (!nx.adj[id].empty()) {
            std::sort(nx.adj[id].begin(), nx.adj[id].end());
            nx.ids.push_back(id);
        }
    return nx;
}

struct Community {
    uint32_t id;
    float weight;
    std::vector<uint32_t> members;
};

}  // namespace

// General path: handles every quirk of louvain.rs (communities that fall apart again, re-keyed vertices).  The flat
// implementation in np2_phase.cpp serves the common case and comes here when it meets a negative community.
std::vector<uint32_t> phase_reads_general(const uint64_t *keys, const long long *vals, uint64_t n_edges, bool asref,
                                          bool use_all_reads) {
    std::map<uint32_t, float> ref_w;
    bool have_ref = false;
    std::set<uint32_t> invalid;
    uint32_t max_id = 0;
    for (uint64_t e = 0; e < n_edges; e++) max_id = std::max(max_id, (uint32_t)keys[e]);  // b > a
    // pass 1: ref pairs (main.rs:972-980)
    for (uint64_t e = 0; e < n_edges; e++) {
        const uint32_t a = (uint32_t)(keys[e] >> 32), b = (uint32_t)keys[e];
        if (a != 0) break;  // keys are sorted: ref pairs come first
        const long long v = vals[e];
        const long long ndif = (v + (1LL << 31)) >> 32;
        const long long sum = v - (ndif << 32);
        if (asref) {
            ref_w[b] = (float)sum;
            have_ref = true;
        }
        if (ndif > 0 && !use_all_reads) invalid.insert(b);
    }
    std::vector<uint8_t> bad_v(max_id + 1, 0), has(max_id + 1, 0);
    for (uint32_t x : invalid) bad_v[x] = 1;
    Level lv;
    lv.grow(max_id);
    for (uint64_t e = 0; e < n_edges; e++) {
        const uint32_t a = (uint32_t)(keys[e] >> 32), b = (uint32_t)keys[e];
        if (a == 0) continue;
        const long long v = vals[e];
        const long long ndif = (v + (1LL << 31)) >> 32;  // number of disagreeing sites
        const long long sum = v - (ndif << 32);          // sum of +-1 over shared heterozygous regions
        const float w = ndif >= 3 ? -(float)ndif : (float)sum;  // main.rs:996-1002
        // main.rs:1004-1010: invalid reads leave the graph, their partners stay (possibly without edges)
        if (!use_all_reads && (bad_v[a] || bad_v[b])) {
            if (!bad_v[a]) has[a] = 1;
            if (!bad_v[b]) has[b] = 1;
            continue;
        }
        has[a] = has[b] = 1;
        lv.adj[a].emplace_back(b, w);  // sorted keys keep every list ascending (smaller neighbours arrive first)
        lv.adj[b].emplace_back(a, w);
    }
    // ---- Louvain (louvain.rs:59-257)
    for (uint32_t v = 0; v <= max_id; v++) {
        if (!has[v]) continue;
        lv.ids.push_back(v);
        lv.cid[v] = v;
        lv.members[v] = {v};
        lv.comm[v] = {v};
    }
    while (move_vertices(lv)) lv = aggregate(lv);
    std::vector<Community> comms;
    for (auto &kv : lv.comm) {  // get_communities louvain.rs:197-245
        if (kv.second.empty()) continue;
        Community c;
        c.id = kv.first;
        c.weight = internal_weight(lv, kv.second);
        for (uint32_t n : kv.second) c.members.insert(c.members.end(), lv.members[n].begin(), lv.members[n].end());
        comms.push_back(std::move(c));
    }
    // weights between communities: sum of the edges from the smaller id's vertices to the larger id's
    std::vector<uint32_t> owner(lv.adj.size(), 0xFFFFFFFFu);
    for (auto &kv : lv.comm)
        for (uint32_t n : kv.second) owner[n] = kv.first;
    std::map<std::pair<uint32_t, uint32_t>, float> between;
    for (auto &kv : lv.comm)
        for (uint32_t n1 : kv.second)
            for (auto &e : lv.adj[n1]) {
                const uint32_t o = owner[e.first];
                if (o != 0xFFFFFFFFu && o > kv.first) between[{kv.first, o}] += e.second;
            }
    std::map<uint32_t, std::set<uint32_t>> conflict;
    for (auto &kv : between) {
        if (kv.second == 0.f) continue;
        if (!(kv.second < 0.f)) herr(NP2_ERR_FORMAT, "the weight of two conflicting community is not less than 0");
        conflict[kv.first.first].insert(kv.first.second);
        conflict[kv.first.second].insert(kv.first.first);
    }
    // ---- phase_communities louvain.rs:290-356
    if (have_ref) {
        std::vector<std::pair<std::pair<int32_t, float>, size_t>> key;
        for (size_t i = 0; i < comms.size(); i++) {
            int32_t cnt = 0;
            float w = 0.f;
            std::vector<uint32_t> mem = comms[i].members;
            std::sort(mem.begin(), mem.end());
            mem.erase(std::unique(mem.begin(), mem.end()), mem.end());
            for (uint32_t n : mem) {
                auto it = ref_w.find(n);
                if (it == ref_w.end()) continue;
                if (it->second > 0.f) cnt++;
                else if (it->second < 0.f) cnt--;
                w += it->second;
            }
            key.push_back({{cnt, w}, i});
        }
        std::stable_sort(key.begin(), key.end(),
                         [](const std::pair<std::pair<int32_t, float>, size_t> &x,
                            const std::pair<std::pair<int32_t, float>, size_t> &y) { return x.first > y.first; });
        std::vector<Community> sorted;
        for (auto &k : key) sorted.push_back(std::move(comms[k.second]));
        comms.swap(sorted);
    } else {
        std::stable_sort(comms.begin(), comms.end(),
                         [](const Community &x, const Community &y) { return x.weight > y.weight; });
    }
    std::set<uint32_t> bad;
    for (size_t p = 0; p < comms.size(); p++) {
        if (bad.count(comms[p].id)) continue;
        auto it = conflict.find(comms[p].id);
        if (it == conflict.end()) continue;
        for (size_t q = p + 1; q < comms.size(); q++)
            if (!bad.count(comms[q].id) && it->second.count(comms[q].id)) bad.insert(comms[q].id);
    }
    std::vector<uint32_t> out(invalid.begin(), invalid.end());
    for (auto &c : comms)
        if (bad.count(c.id)) out.insert(out.end(), c.members.begin(), c.members.end());
    std::sort(out.begin(), out.end());
    out.erase(std::unique(out.begin(), out.end()), out.end());
    return out;
}

/* ================================================================= consensus patching */

namespace {
template <class F>
void for_each_choice(const std::vector<uint32_t> &lens, F f) {  // itertools::multi_cartesian_product order
    std::vector<uint32_t> ch(lens.size(), 0);
    for (auto l : lens)
        if (!l) return;
    for (;;) {
        f(ch);
        size_t d = lens.size();
        while (d-- > 0) {
            if (++ch[d] < lens[d]) break;
            ch[d] = 0;
        }
        if (d == (size_t)-1) break;
    }
}
// the `need` bases that precede region q in the current (patched) consensus, in forward order
void take_left(const Patched &pc, size_t q, size_t need, std::vector<uint8_t> &out) {
    std::vector<uint8_t> rev;
    uint64_t i = pc.a[q];  // exclusive end of the DP stretch before the region
    size_t rq = q;
    while (rev.size() < need) {
        const uint64_t lo = rq > 0 ? pc.b[rq - 1] : 0;
        while (i > lo && rev.size() < need) rev.push_back(pc.cbase[--i]);
        if (rev.size() >= need || rq == 0) break;
        rq--;
        const Allele &al = pc.seed[rq];
        for (uint32_t x = al.len; x-- > 0 && rev.size() < need;) rev.push_back(al.s[x]);
        i = pc.a[rq];
    }
    out.insert(out.end(), rev.rbegin(), rev.rend());
}
void take_right(const Patched &pc, size_t q, size_t need, std::vector<uint8_t> &out) {
    size_t got = 0;
    uint64_t i = pc.b[q];
    size_t rq = q;
    const size_t nr = pc.a.size();
    while (got < need) {
        const uint64_t hi = rq + 1 < nr ? pc.a[rq + 1] : pc.N;
        while (i < hi && got < need) {
            out.push_back(pc.cbase[i++]);
            got++;
        }
        if (got >= need || rq + 1 >= nr) break;
        rq++;
        const Allele &al = pc.seed[rq];
        for (uint32_t x = 0; x < al.len && got < need; x++) {
            out.push_back(al.s[x]);
            got++;
        }
        i = pc.b[rq];
    }
}
}  // namespace

void reupdate_build(const Patched &pc, uint32_t k, Reupdate &ru) {
    ru = Reupdate();
    for (size_t q = 0; q < pc.lable.size(); q++)
        if (pc.lable[q] & LABLE_RECH) ru.rech.push_back((uint32_t)q);
    ru.off.push_back(0);
    auto put = [&](const Allele &al) { ru.pool.insert(ru.pool.end(), al.s, al.s + al.len); };
    std::vector<uint8_t> left, right;
    size_t sj = 0;
    while (sj < ru.rech.size()) {
        size_t ej = sj + 1;  // chain regions closer than k, at most 6 per group (main.rs:1197-1206)
        while (ej < ru.rech.size() && pc.start[ru.rech[ej]] < pc.end[ru.rech[ej - 1]] + k) {
            ej++;
            if (ej > sj + 5) break;
        }
        // flanks: k-1 bases of the current consensus on either side (iter_consensus_extend main.rs:1100-1139)
        left.clear();
        right.clear();
        take_left(pc, ru.rech[sj], k - 1, left);
        take_right(pc, ru.rech[ej - 1], k - 1, right);
        Reupdate::Group g{(uint32_t)sj, (uint32_t)ej, ru.off.size() - 1};
        if (ej == sj + 1) {
            for (const Allele &al : pc.cand[ru.rech[sj]]) {
                ru.pool.insert(ru.pool.end(), left.begin(), left.end());
                put(al);
                ru.pool.insert(ru.pool.end(), right.begin(), right.end());
                ru.off.push_back(ru.pool.size());
            }
        } else {
            std::vector<uint32_t> lens;
            for (size_t x = sj; x < ej; x++) lens.push_back((uint32_t)pc.cand[ru.rech[x]].size());
            for_each_choice(lens, [&](const std::vector<uint32_t> &ch) {
                ru.pool.insert(ru.pool.end(), left.begin(), left.end());
                for (size_t x = 0; x < ch.size(); x++) {
                    const uint32_t q = ru.rech[sj + x];
                    put(pc.cand[q][ch[x]]);
                    if (x + 1 < ch.size()) {
                        // consensus strictly between the two regions (iter_consensus_region main.rs:1068-1097):
                        // the DP stretch between their index ranges, plus any non-RECH regions' alleles inside it
                        const uint32_t qn = ru.rech[sj + x + 1];
                        uint64_t i = pc.b[q];
                        for (uint32_t m = q + 1; m <= qn; m++) {
                            ru.pool.insert(ru.pool.end(), pc.cbase + i, pc.cbase + pc.a[m]);
                            if (m < qn) {
                                ru.pool.insert(ru.pool.end(), pc.seed[m].s, pc.seed[m].s + pc.seed[m].len);
                                i = pc.b[m];
                            }
                        }
                    }
                }
                ru.pool.insert(ru.pool.end(), right.begin(), right.end());
                ru.off.push_back(ru.pool.size());
                if (ru.pool.size() > (1ull << 32)) herr(NP2_ERR_UNSUPPORTED, "cartesian re-check group too large");
            });
        }
        ru.groups.push_back(g);
        sj = ej;
    }
}

void reupdate_apply(Patched &pc, const Reupdate &ru, const uint16_t *ks, uint32_t iter_count) {
    for (auto &g : ru.groups) {
        uint64_t si = g.first_string;
        if (g.ej == g.sj + 1) {
            for (Allele &al : pc.cand[ru.rech[g.sj]]) al.kscore = ks[si++];
        } else {
            std::vector<uint32_t> lens;
            for (size_t x = g.sj; x < g.ej; x++) lens.push_back((uint32_t)pc.cand[ru.rech[x]].size());
            for (size_t x = g.sj; x < g.ej; x++)
                for (Allele &al : pc.cand[ru.rech[x]]) al.kscore = 0;
            // later combinations overwrite earlier ones (main.rs:1351-1366)
            for_each_choice(lens, [&](const std::vector<uint32_t> &ch) {
                const uint16_t s = ks[si++];
                if (s > 0)
                    for (size_t x = 0; x < ch.size(); x++) pc.cand[ru.rech[g.sj + x]][ch[x]].kscore = s;
            });
        }
    }
    // choose the allele (main.rs:1371-1406), then the lable bookkeeping (main.rs:1411-1417)
    for (uint32_t q : ru.rech) {
        auto &cd = pc.cand[q];
        size_t c = 0, valid = 0;
        for (size_t p = 0; p < cd.size(); p++)
            if (cd[p].kscore != 0) {
                if (c == 0 || cd[p].order == 0) c = p + 1;
                valid++;
            }
        if (c != 0) {
            pc.seed[q] = cd[c - 1];
        } else if (iter_count == 1) {
            size_t i = 0;
            for (size_t p = 0; p < cd.size(); p++)
                if (cd[p].order == 0) {
                    i = p;
                    break;
                }
            pc.seed[q] = cd[i];
        }
        if (valid <= 1) pc.lable[q] ^= LABLE_RECH;
    }
}

void positions(const Patched &pc, std::vector<uint32_t> &pos) {
    const size_t nr = pc.a.size();
    uint64_t total = pc.N;
    for (size_t q = 0; q < nr; q++) total = total - (pc.b[q] - pc.a[q]) + pc.seed[q].len;
    pos.resize(total);
    uint64_t w = 0, i = 0;
    for (size_t q = 0; q <= nr; q++) {
        const uint64_t hi = q < nr ? pc.a[q] : pc.N;
        if (hi > i) {
            memcpy(pos.data() + w, pc.cpos + i, (hi - i) * 4);
            w += hi - i;
        }
        if (q == nr) break;
        std::fill(pos.begin() + w, pos.begin() + w + pc.seed[q].len, pc.start[q]);
        w += pc.seed[q].len;
        i = pc.b[q];
    }
}

}  // namespace np2
