/*
 * np2_oracle.cpp — CPU restatement of the NextPolish2 per-contig polish path.
 *
 * TEST INFRASTRUCTURE ONLY (see np2_oracle.h).  Every function cites the
 * reference lines it follows (paths relative to /root/reference).  The
 * restatement keeps the reference's data structures (per-position Vec of
 * 3-mers, String candidates, nested maps) on purpose: it is the slow,
 * obviously-faithful version the CUDA path is checked against.
 *
 * PARITY UNPINNED for the Rust path (no cargo/rustc here, reference has no
 * tests); hash/lookup is pinned against compiled yak (oracle/_ref).
 *
 * Known, documented divergence risk: FxHashMap iteration order leaks into
 * louvain.rs:123,145-165,199; we iterate in ascending-id order (SURVEY §7
 * hard part 3).
 */
#include "np2_oracle.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

namespace {

thread_local std::string g_err;

struct OracleError : std::runtime_error {
    using std::runtime_error::runtime_error;
};
[[noreturn]] void fail(const std::string &m) { throw OracleError(m); }

/* ------------------------------------------------------------------ kmer.rs */

// kmer.rs:11-22
const uint8_t SEQ_NUM[128] = {
    65, 67, 71, 84, 45, 78, 77, 4, 4, 4, 4, 4, 4, 4, 4, 4,
    4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
    4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
    4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
    4, 0, 4, 1, 4, 4, 4, 2, 4, 4, 4, 4, 4, 6, 5, 4,
    4, 4, 4, 4, 3, 3, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
    4, 0, 4, 1, 4, 4, 4, 2, 4, 4, 4, 4, 4, 6, 5, 4,
    4, 4, 4, 4, 3, 3, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
};
inline uint8_t seq_num(uint8_t c) {
    if (c >= 128) fail("non-ASCII byte indexes SEQ_NUM out of range (reference would panic)");
    return SEQ_NUM[c];
}

// kmer.rs:223-233 (yak/yak-priv.h:10)
uint64_t yak_hash64(uint64_t key, uint64_t mask) {
    key = (~key + (key << 21)) & mask;
    key = key ^ key >> 24;
    key = ((key + (key << 3)) + (key << 8)) & mask;
    key = key ^ key >> 14;
    key = ((key + (key << 2)) + (key << 4)) & mask;
    key = key ^ key >> 28;
    key = (key + (key << 31)) & mask;
    return key;
}
// kmer.rs:235-244
uint64_t yak_hash64_64(uint64_t key) {
    key = ~key + (key << 21);
    key = key ^ key >> 24;
    key = (key + (key << 3)) + (key << 8);
    key = key ^ key >> 14;
    key = (key + (key << 2)) + (key << 4);
    key = key ^ key >> 28;
    key = key + (key << 31);
    return key;
}
// kmer.rs:246-249
uint64_t yak_hash_long(const uint64_t x[4]) {
    int j = x[1] < x[3] ? 0 : 1;
    return yak_hash64_64(x[j << 1]) + yak_hash64_64(x[j << 1 | 1]);
}

// kmer.rs:255-314: calls f(kmer) for every canonical k-mer of the byte stream.
// For ksize < 32 the value is the un-hashed canonical 2-bit k-mer, for ksize >= 32
// it is already yak_hash_long (kmer.rs:306-308).
template <class It, class F>
void iter2kmer(It it, It end, size_t ksize, F f) {
    size_t l = 0;
    if (ksize < 32) {
        const uint64_t shift = 2 * ((uint64_t)ksize - 1);
        const uint64_t mask = (1ULL << (2 * (uint64_t)ksize)) - 1;
        uint64_t kmer[2] = {0, 0};
        for (; it != end; ++it) {
            uint64_t c = seq_num(*it);
            if (c < 4) {
                kmer[0] = (kmer[0] << 2 | c) & mask;
                kmer[1] = (kmer[1] >> 2) | (3 ^ c) << shift;
                l += 1;
            } else {
                l = 0;
            }
            if (l >= ksize) f(kmer[0] < kmer[1] ? kmer[0] : kmer[1]);
        }
    } else {
        const uint64_t shift = (uint64_t)ksize - 1;
        const uint64_t mask = (1ULL << (uint64_t)ksize) - 1;
        uint64_t kmer[4] = {0, 0, 0, 0};
        for (; it != end; ++it) {
            uint64_t c = seq_num(*it);
            if (c < 4) {
                kmer[0] = (kmer[0] << 1 | (c & 1)) & mask;
                kmer[1] = (kmer[1] << 1 | (c >> 1)) & mask;
                kmer[2] = kmer[2] >> 1 | (1 - (c & 1)) << shift;
                kmer[3] = kmer[3] >> 1 | (1 - (c >> 1)) << shift;
                l += 1;
            } else {
                l = 0;
                kmer[0] = kmer[1] = kmer[2] = kmer[3] = 0;
            }
            if (l >= ksize) f(yak_hash_long(kmer));
        }
    }
}

}  // namespace

/* ---------------------------------------------------------- yak tables */

// In-memory image of a yak dump (yak/htab.c:190-211) + the query protocol of
// KmerInfo (kmer.rs:62-221): clear / insert / retrieve_kmers / get.
struct np2o_table {
    uint32_t k = 0, pre = 10;
    uint64_t n = 0;
    // open addressing on the full hash h (bucket id = h & 1023 restored from the file layout)
    std::vector<uint64_t> slot_hash;
    std::vector<uint16_t> slot_cnt1;  // count + 1, 0 = empty
    uint64_t cap_mask = 0;
    bool stream_scan = false;

    void reserve(uint64_t nkeys) {
        uint64_t cap = 16;
        while (cap < nkeys * 2) cap <<= 1;
        slot_hash.assign(cap, 0);
        slot_cnt1.assign(cap, 0);
        cap_mask = cap - 1;
    }
    static inline uint64_t mix(uint64_t h) {
        h ^= h >> 33;
        h *= 0xff51afd7ed558ccdULL;
        h ^= h >> 33;
        return h;
    }
    void put(uint64_t h, uint16_t c) {
        uint64_t i = mix(h) & cap_mask;
        while (slot_cnt1[i] != 0) {
            if (slot_hash[i] == h) {
                slot_cnt1[i] = c + 1;
                return;
            }
            i = (i + 1) & cap_mask;
        }
        slot_hash[i] = h;
        slot_cnt1[i] = c + 1;
        n++;
    }
    // -1 = absent
    int find(uint64_t h) const {
        uint64_t i = mix(h) & cap_mask;
        while (slot_cnt1[i] != 0) {
            if (slot_hash[i] == h) return (int)slot_cnt1[i] - 1;
            i = (i + 1) & cap_mask;
        }
        return -1;
    }
};

namespace {

// KmerInfo query state (kmer.rs:62-69): one per job and table, like the per-thread clone (main.rs:1724).
struct KmerInfo {
    const np2o_table *tab;
    uint32_t ksize;
    uint64_t kmask;
    std::unordered_map<uint64_t, uint16_t> q;  // inserted hashes -> count (0 until retrieved)

    explicit KmerInfo(const np2o_table *t) : tab(t), ksize(t->k) {
        kmask = ksize < 32 ? (1ULL << (2 * (uint64_t)ksize)) - 1 : 0;
    }
    // kmer.rs:102-110
    uint64_t to_hash(uint64_t kmer) const { return ksize < 32 ? yak_hash64(kmer, kmask) : kmer; }
    void clear() { q.clear(); }                   // kmer.rs:216
    void insert(uint64_t hash) { q.emplace(hash, 0); }  // kmer.rs:113 (clear_count = true)
    // kmer.rs:132-170: keep the stored 10-bit count of every queried key whose count >= min_count
    void retrieve_kmers(uint16_t min_count) {
        if (tab->stream_scan) {
            // the reference's way: stream every key of the table past the query set
            const size_t cap = tab->slot_hash.size();
            for (size_t i = 0; i < cap; i++) {
                if (!tab->slot_cnt1[i]) continue;
                uint16_t c = tab->slot_cnt1[i] - 1;
                if (c < min_count) continue;
                auto it = q.find(tab->slot_hash[i]);
                if (it != q.end()) it->second = c;
            }
        } else {
            for (auto &kv : q) {
                int c = tab->find(kv.first);
                if (c >= 0 && (uint16_t)c >= min_count) kv.second = (uint16_t)c;
            }
        }
    }
    // kmer.rs:123-125; None (never inserted) reported as -1
    int get(uint64_t hash) const {
        auto it = q.find(hash);
        return it == q.end() ? -1 : (int)it->second;
    }
};

/* ------------------------------------------------------------------ main.rs */

const size_t LQSEQ_MAX_CAN_COUNT = 60;       // main.rs:30
const uint64_t INVALID_KMER = UINT64_MAX;    // main.rs:31

// main.rs:33-52
struct AlignBase {
    uint8_t q_base = 0;
    uint16_t delta = 0;
    uint32_t t_pos = 0;
    static AlignBase head(uint32_t t_pos, uint16_t delta) {
        AlignBase b;
        b.q_base = 0b1111;
        b.delta = delta;
        b.t_pos = t_pos;
        return b;
    }
    bool is_head() const { return q_base == 0b1111; }
    bool operator==(const AlignBase &o) const { return q_base == o.q_base && delta == o.delta && t_pos == o.t_pos; }
};

// main.rs:55-185
struct Kmer {
    uint16_t delta = 0;
    uint16_t bases = 0;
    uint32_t count = 0;
    mutable uint32_t besti = 0;
    mutable int64_t score = 0;

    static Kmer make(const AlignBase &b1, const AlignBase &b2, const AlignBase &b3) {  // main.rs:84-102
        uint16_t bases = 0;
        if (b2.t_pos == b1.t_pos) bases |= 0b0100;
        if (b2.t_pos == b3.t_pos) bases |= 0b0001;
        Kmer k;
        k.delta = b1.delta;
        k.bases = (uint16_t)((((uint16_t)(bases << 4 | b1.q_base) << 4 | b2.q_base) << 4) | b3.q_base);
        k.count = 1;
        return k;
    }
    // main.rs:105-184; u32 / u16 arithmetic wraps (release build, no overflow checks)
    void get_bases(uint32_t p, AlignBase &a, AlignBase &b, AlignBase &c) const {
        uint8_t q1 = bases >> 8 & 0b1111, q2 = bases >> 4 & 0b1111, q3 = bases & 0b1111;
        if ((bases & 0b0101000000000000) == 0b0101000000000000) {  // A--
            a = {q1, delta, p};
            b = {q2, (uint16_t)(delta + 1), p};
            c = {q3, (uint16_t)(delta + 2), p};
        } else if ((bases & 0b0001000000000000) != 0) {  // AA-
            a = {q1, delta, p - 1};
            b = {q2, 0, p};
            c = {q3, 1, p};
        } else if ((bases & 0b0100000000000000) != 0) {  // A-A
            a = {q1, delta, p - 1};
            b = {q2, (uint16_t)(delta + 1), p - 1};
            c = {q3, 0, p};
        } else {  // AAA
            a = {q1, delta, p - 2};
            b = {q2, 0, p - 1};
            c = {q3, 0, p};
        }
    }
    uint16_t b3_delta() const {
        AlignBase a, b, c;
        get_bases(0, a, b, c);
        return c.delta;
    }
};

// main.rs:187-250
struct Msa {
    std::vector<Kmer> kmers;
    void push(const Kmer &v) {  // main.rs:193-207
        for (auto &k : kmers) {
            if (k.bases == v.bases && k.delta == v.delta) {
                k.count += 1;
                if (!(k.count < UINT32_MAX)) fail("kmer count overflow!");
                return;
            }
        }
        kmers.push_back(v);
    }
    void sort() {  // main.rs:227-229 (sort_by_cached_key is stable)
        std::stable_sort(kmers.begin(), kmers.end(),
                         [](const Kmer &x, const Kmer &y) { return x.b3_delta() < y.b3_delta(); });
    }
    int64_t coverage() const {  // main.rs:232-241
        int64_t c = 0;
        for (auto &k : kmers) {
            if (k.b3_delta() != 0) break;
            c += (int64_t)k.count;
        }
        return c;
    }
};

const uint32_t ALN_T_S_LABLE = 1u << 31;  // main.rs:271

// main.rs:353-529
struct Alignment {
    uint32_t shift = 0, aln_t_s = 0, aln_t_e = 0, aln_q_s = 0, aln_q_e = 0;
    std::string q_aln_str, t_aln_str;

    void clear() {
        shift = aln_t_s = aln_t_e = aln_q_s = aln_q_e = 0;
        q_aln_str.clear();
        t_aln_str.clear();
    }
    size_t aln_len() const { return t_aln_str.size() - shift; }  // main.rs:442-444

    // main.rs:386-440.  cigar: BAM-encoded ops (len<<4|op), tseq: reference from aln_t_s on,
    // seq4: BAM 4-bit SEQ, l_seq its length in bases.
    void fill_with_cigar(const uint32_t *cigar, uint32_t n_cigar, const uint8_t *tseq, uint64_t tlen_left,
                         const uint8_t *seq4, uint32_t l_seq) {
        static const char DEC[] = "=ACMGRSVTWYHKDBN";  // rust-htslib Seq index decode
        uint32_t qs = 0, ts = 0;
        bool is_first = true;
        for (uint32_t ci = 0; ci < n_cigar; ci++) {
            uint32_t l = cigar[ci] >> 4, op = cigar[ci] & 15;
            switch (op) {
                case 4:  // S
                    qs += l;
                    if (is_first) aln_q_s = qs;
                    else aln_q_e = qs - l;
                    break;
                case 0: case 7: case 8:  // M = X
                    if ((uint64_t)qs + l > l_seq) fail("CIGAR consumes more query bases than SEQ holds");
                    if ((uint64_t)ts + l > tlen_left) fail("alignment runs past the end of the contig");
                    for (uint32_t i = 0; i < l; i++) {
                        q_aln_str.push_back(DEC[(seq4[qs >> 1] >> ((~qs & 1) << 2)) & 15]);
                        qs += 1;
                    }
                    t_aln_str.append((const char *)tseq + ts, l);
                    ts += l;
                    break;
                case 1:  // I
                    if ((uint64_t)qs + l > l_seq) fail("CIGAR consumes more query bases than SEQ holds");
                    for (uint32_t i = 0; i < l; i++) {
                        q_aln_str.push_back(DEC[(seq4[qs >> 1] >> ((~qs & 1) << 2)) & 15]);
                        qs += 1;
                    }
                    t_aln_str.append(l, '-');
                    break;
                case 2:  // D
                    if ((uint64_t)ts + l > tlen_left) fail("alignment runs past the end of the contig");
                    q_aln_str.append(l, '-');
                    t_aln_str.append((const char *)tseq + ts, l);
                    ts += l;
                    break;
                case 5:  // H
                    break;
                default:
                    fail("Unknown cigar");
            }
            is_first = false;
        }
        if (aln_q_e == 0) aln_q_e = qs;
        aln_t_e = aln_t_s + ts;
    }

    // main.rs:447-513
    void trim(uint32_t len) {
        uint32_t j = 0;
        const size_t n = t_aln_str.size();
        const uint8_t *t = (const uint8_t *)t_aln_str.data();
        const uint8_t *q = (const uint8_t *)q_aln_str.data();
        for (size_t i = 0; i < n; i++) {
            if (t[i] == q[i]) {
                j += 1;
                aln_t_s += 1;
                aln_q_s += 1;
            } else {
                if (t[i] != '-') aln_t_s += 1;
                if (q[i] != '-') aln_q_s += 1;
                j = 0;
            }
            if (j == len) {
                aln_t_s -= len;
                aln_q_s -= len;
                shift = (uint32_t)i + 1 - len;
                break;
            }
        }
        if (j == len) {
            j = 0;
            for (size_t i = n; i-- > 0;) {
                if (t[i] == q[i]) {
                    j += 1;
                    aln_t_e -= 1;
                    aln_q_e -= 1;
                } else {
                    if (t[i] != '-') aln_t_e -= 1;
                    if (q[i] != '-') aln_q_e -= 1;
                    j = 0;
                }
                if (j == len) {
                    aln_t_e += len;
                    aln_q_e += len;
                    size_t new_len = i + len;
                    if (new_len < n) {
                        t_aln_str.resize(new_len);
                        q_aln_str.resize(new_len);
                    }
                    break;
                }
            }
        } else {
            shift = (uint32_t)n;
        }
    }
};

// main.rs:272-351
struct AlignSeq {
    uint32_t aln_t_s = 0, aln_t_e = 0;
    std::vector<uint8_t> align_bases;

    static AlignSeq make(const Alignment &aln) {  // main.rs:279-312
        size_t len = (aln.aln_len() + 1) >> 1;
        AlignSeq a;
        a.aln_t_s = aln.aln_t_s;
        a.aln_t_e = aln.aln_t_s;
        a.align_bases.assign(len + 1, 0);
        size_t i = 0;
        for (size_t c = aln.shift; c < aln.t_aln_str.size(); c++) {
            uint8_t tb = (uint8_t)aln.t_aln_str[c], qb = (uint8_t)aln.q_aln_str[c];
            uint8_t b = seq_num(qb);
            if (tb == '-') b |= 8;
            else if (i != 0) a.aln_t_e += 1;
            if ((i & 1) == 0) b <<= 4;
            a.align_bases[i >> 1] |= b;
            i += 1;
        }
        if ((i & 1) == 0) a.align_bases[i >> 1] |= 255;
        else a.align_bases[i >> 1] |= 15;
        return a;
    }
    // main.rs:314-338
    bool get_align_tag(size_t &p, AlignBase &ab) const {
        uint8_t t = align_bases[p >> 1];
        if ((p & 1) == 0) t >>= 4;
        if ((t & 15) == 15) return false;
        ab.q_base = t & 7;
        if (p != 0) {
            if (t & 8) ab.delta += 1;
            else {
                ab.delta = 0;
                ab.t_pos += 1;
            }
        } else {
            ab.t_pos = aln_t_s;
            ab.delta = 0;
        }
        p += 1;
        return true;
    }
    void set_lable() { aln_t_s |= ALN_T_S_LABLE; }
    bool has_lable() const { return (aln_t_s & ALN_T_S_LABLE) != 0; }
    void unset_lable() { aln_t_s ^= ALN_T_S_LABLE; }
};

// main.rs:531-574
void filter_alignseqs_by_clip(std::vector<AlignSeq> &alignseqs) {
    const uint32_t offset = 50;
    std::vector<std::pair<uint32_t, uint32_t>> aln_ranges;
    uint32_t s = 0, e = 0;
    for (auto &x : alignseqs) {
        if (x.has_lable()) continue;
        uint32_t aln_t_s = x.aln_t_s + offset, aln_t_e = x.aln_t_e - offset;
        if (s == e) {
            s = aln_t_s;
            e = aln_t_e;
        } else if (aln_t_s > e) {
            aln_ranges.emplace_back(s, e);
            s = aln_t_s;
            e = aln_t_e;
        } else if (e < aln_t_e) {
            e = aln_t_e;
        }
    }
    if (s != e) aln_ranges.emplace_back(s, e);
    for (auto &a : alignseqs) {
        if (!a.has_lable()) continue;
        a.unset_lable();
        bool in = false;
        for (auto &r : aln_ranges) {
            if (r.first <= a.aln_t_s && a.aln_t_e <= r.second) {
                in = true;
                break;
            } else if (a.aln_t_e < r.first) {
                break;
            }
        }
        if (in) a.align_bases.clear();
    }
}

// main.rs:576-589
void update_msas(std::vector<Msa> &msas, const std::vector<AlignSeq> &alignseqs) {
    for (auto &alignseq : alignseqs) {
        if (alignseq.align_bases.empty()) continue;
        size_t p = 0;
        AlignBase b1 = AlignBase::head(alignseq.aln_t_s - 1, 0);
        AlignBase b2 = AlignBase::head(alignseq.aln_t_s - 1, 1);
        AlignBase b3;
        while (alignseq.get_align_tag(p, b3)) {
            if (b3.t_pos >= msas.size()) fail("alignment column beyond contig end");
            msas[b3.t_pos].push(Kmer::make(b1, b2, b3));
            b1 = b2;
            b2 = b3;
        }
    }
}

struct ConsensusBase {  // main.rs:591-596
    uint32_t pos;
    uint8_t base;
};

struct LqSeq {  // main.rs:647-653
    uint32_t order = 0;
    uint16_t kscore = 0;
    uint64_t kmer = 0;
    std::string seq;
};

const uint8_t LQSEQS_LABLE_TEMP = 0b00000001;  // main.rs:655-658
const uint8_t LQSEQS_LABLE_SUCC = 0b10000000;
const uint8_t LQSEQS_LABLE_HETE = 0b01000000;
const uint8_t LQSEQS_LABLE_RECH = 0b00100000;

struct LqSeqs {  // main.rs:660-727
    uint8_t lable = 0;
    uint32_t start = 0, end = 0;
    std::string sudoseed;
    std::vector<LqSeq> seqs;
    void set_lable(uint8_t l) { lable |= l; }
    void unset_lable(uint8_t l) { lable ^= l; }
    bool has_lable(uint8_t l) const { return (lable & l) != 0; }
    void clean_seqs() { std::vector<LqSeq>().swap(seqs); }
    void retain_sort_seqs(const std::map<uint32_t, size_t> &stat, size_t min_c) {  // main.rs:714-726
        auto get = [&](uint32_t order) -> size_t {
            auto it = stat.find(order);
            return it == stat.end() ? 0 : it->second;
        };
        std::stable_sort(seqs.begin(), seqs.end(),
                         [&](const LqSeq &a, const LqSeq &b) { return get(a.order) > get(b.order); });
        size_t c = 0;
        for (auto &v : seqs) {
            if (get(v.order) < min_c) break;
            c += 1;
        }
        seqs.resize(c);
    }
};

// main.rs:740-778
void retrieve_kmer_count(std::vector<LqSeqs> &lqseqs, KmerInfo &kmer_info, uint16_t min_kmer_count) {
    kmer_info.clear();
    const size_t ksize = kmer_info.ksize;
    for (auto &lqseq : lqseqs)
        for (auto &seq : lqseq.seqs) {
            if (seq.seq.size() > ksize) {
                iter2kmer((const uint8_t *)seq.seq.data(), (const uint8_t *)seq.seq.data() + seq.seq.size(), ksize,
                          [&](uint64_t kmer) { kmer_info.insert(kmer_info.to_hash(kmer)); });
            } else if (seq.kmer != INVALID_KMER) {
                kmer_info.insert(seq.kmer);
            }
        }
    kmer_info.retrieve_kmers(min_kmer_count);
    for (auto &lqseq : lqseqs)
        for (auto &seq : lqseq.seqs) {
            if (seq.seq.size() > ksize) {
                bool any = false;
                uint16_t mn = 0;
                iter2kmer((const uint8_t *)seq.seq.data(), (const uint8_t *)seq.seq.data() + seq.seq.size(), ksize,
                          [&](uint64_t x) {
                              int g = kmer_info.get(kmer_info.to_hash(x));
                              uint16_t v = g < 0 ? 0 : (uint16_t)g;
                              if (!any || v < mn) mn = v;
                              any = true;
                          });
                seq.kscore = any ? mn : 0;
            } else if (seq.kmer != INVALID_KMER) {
                int g = kmer_info.get(seq.kmer);
                seq.kscore = g < 0 ? 0 : (uint16_t)g;
            }
        }
}

// main.rs:780-801
bool is_valid_snp(const std::string &seq1, const std::string &seq2) {
    size_t i = 0, j = 0;
    while (i < seq1.size() && j < seq2.size()) {
        if (seq1[i] != seq2[j]) return true;
        while (i + 1 < seq1.size() && seq1[i] == seq1[i + 1]) i += 1;
        while (j + 1 < seq2.size() && seq2[j] == seq2[j + 1]) j += 1;
        i += 1;
        j += 1;
    }
    return false;
}

// main.rs:803-811
size_t get_min_count(size_t c) { return c >= 9 ? 3 : (c >= 6 ? 2 : 1); }

struct OrderStat {
    size_t max1_c, max1_p, max2_c, max2_p;
};
// main.rs:813-849
OrderStat fill_order_stat(const LqSeqs &lqseq, size_t *stats, std::map<uint32_t, size_t> &order_stat) {
    size_t max1_c = 0, max1_p = 0, max2_c = 0, max2_p = 0;
    std::fill(stats, stats + LQSEQ_MAX_CAN_COUNT, 0);
    order_stat.clear();
    const size_t n = lqseq.seqs.size();
    for (size_t p1 = 0; p1 < n; p1++) {
        const LqSeq &seq = lqseq.seqs[p1];
        if (!(seq.kscore > 0)) continue;
        if (stats[p1] > 0) continue;
        size_t c = 0;
        for (size_t x = p1; x < n; x++)
            if (lqseq.seqs[x].seq == seq.seq) c++;
        order_stat[lqseq.seqs[p1].order] = c;
        for (size_t x = p1; x < n; x++)
            if (lqseq.seqs[x].seq == seq.seq) stats[x] = c;
        if (c > max1_c || (c == max1_c && seq.order == 0)) {
            max2_c = max1_c;
            max2_p = max1_p;
            max1_c = c;
            max1_p = p1;
        } else if (max1_p == max2_p || c > max2_c) {
            max2_c = c;
            max2_p = p1;
        }
    }
    return {max1_c, max1_p, max2_c, max2_p};
}

// main.rs:851-860
bool no_dupseq_lqseq(const LqSeqs &lqseq) {
    for (size_t p1 = 1; p1 < lqseq.seqs.size(); p1++)
        for (size_t p2 = p1 + 1; p2 < lqseq.seqs.size(); p2++)
            if (lqseq.seqs[p1].seq == lqseq.seqs[p2].seq) return false;
    return true;
}

// main.rs:862-914
void fill_seed_lqseqs(std::vector<LqSeqs> &lqseqs, long max_indel_len) {
    size_t stats[LQSEQ_MAX_CAN_COUNT];
    std::map<uint32_t, size_t> order_stat;
    for (auto &lqseq : lqseqs) {
        if (lqseq.seqs.empty()) fail("LQ region without any candidate (reference would panic)");
        OrderStat os = fill_order_stat(lqseq, stats, order_stat);
        size_t max1_c = os.max1_c, max1_p = os.max1_p;
        lqseq.sudoseed = lqseq.seqs[max1_p].seq;
        lqseq.set_lable(LQSEQS_LABLE_SUCC);
        lqseq.set_lable(LQSEQS_LABLE_RECH);
        size_t min_c = get_min_count(lqseq.seqs.size());
        if (lqseq.seqs[0].order != 0) fail("the first lqseq is not ref.");
        auto it0 = order_stat.find(0);
        if (it0 != order_stat.end()) {
            if (it0->second > 1 && it0->second < min_c) it0->second = min_c;
        } else {
            size_t c = 0;
            for (auto &x : lqseq.seqs)
                if (x.seq == lqseq.seqs[0].seq) c++;
            if (c > 1) order_stat[0] = min_c;
        }
        if (max1_p != 0 && max1_c < min_c && (max1_c > 1 || no_dupseq_lqseq(lqseq))) {
            order_stat.at(lqseq.seqs[max1_p].order) = min_c;
            order_stat[0] = min_c;
        } else if (max1_c < min_c) {
            order_stat[0] = min_c;
        }
        lqseq.retain_sort_seqs(order_stat, min_c);
        if (lqseq.seqs.empty()) fail("no candidate survives retain_sort_seqs (reference would panic)");
        bool skip_long_lqseq =
            std::labs((long)lqseq.sudoseed.size() - (long)lqseq.seqs[0].seq.size()) > max_indel_len;
        if (lqseq.seqs.size() <= 1 || skip_long_lqseq) {
            if (!lqseq.seqs.empty() || skip_long_lqseq) lqseq.sudoseed = lqseq.seqs[0].seq;
            lqseq.unset_lable(LQSEQS_LABLE_RECH);
            lqseq.clean_seqs();
        }
    }
}

// main.rs:916-946
void mark_hete_lqseqs(std::vector<LqSeqs> &lqseqs) {
    size_t stats[LQSEQ_MAX_CAN_COUNT];
    std::map<uint32_t, size_t> order_stat;
    for (auto &lqseq : lqseqs) {
        OrderStat os = fill_order_stat(lqseq, stats, order_stat);
        size_t min_c = get_min_count(lqseq.seqs.size());
        if (lqseq.seqs.empty()) continue;  // max2_c = 0 < min_c in the reference
        if (os.max2_c >= min_c &&
            (lqseq.seqs[os.max1_p].seq.size() == lqseq.seqs[os.max2_p].seq.size() ||
             (lqseq.seqs.size() >= 6 && os.max2_c >= os.max1_c / 2)) &&
            is_valid_snp(lqseq.seqs[os.max1_p].seq, lqseq.seqs[os.max2_p].seq)) {
            lqseq.set_lable(LQSEQS_LABLE_HETE);
            for (size_t p = 0; p < lqseq.seqs.size(); p++) {
                if (!(lqseq.seqs[p].kscore > 0)) continue;
                if (stats[p] < min_c) lqseq.seqs[p].kscore = 0;
            }
        }
    }
}

/* ------------------------------------------------------------- louvain.rs */

// Where the reference's result can depend on FxHashMap iteration order (SURVEY hard part 3), counted so that the
// exposure can be MEASURED on real inputs (np2o_order_exposure):
//  [0] calls of phase_communities, [1] communities declustered in second_stage (louvain.rs:136-165: their new ids depend
//  on the order communities and nodes are visited in), [2] pairs of communities with EQUAL sort keys that conflict with
//  each other while both are still valid (louvain.rs:316-339: which of the two is dropped depends on their order
//  before the stable sort), [3] reads in the communities of [2].
static std::atomic<uint64_t> g_order_exposure[4];

typedef std::map<uint32_t, std::map<uint32_t, float>> Graph;  // louvain.rs:31 (ascending-id iteration, see header)

void insert_data(Graph &data, uint32_t k1, uint32_t k2, float v) {  // louvain.rs:273-279
    auto it = data.find(k1);
    if (it != data.end()) {
        auto jt = it->second.find(k2);
        if (jt != it->second.end()) jt->second += v;
        else it->second.emplace(k2, v);
    } else {
        data[k1].emplace(k2, v);
    }
}
void assign_data(Graph &data, uint32_t k1, uint32_t k2, float v) {  // louvain.rs:282-288
    data[k1][k2] = v;
}

struct Node {  // louvain.rs:13-17
    uint32_t id = 0;
    float weight = 0;
    std::set<uint32_t> nodes;
};

struct Louvain {  // louvain.rs:30-257
    Graph data;
    std::map<uint32_t, std::set<uint32_t>> communities;
    std::map<uint32_t, Node> node;

    explicit Louvain(Graph d) : data(std::move(d)) {  // louvain.rs:60-70
        for (auto &kv : data) {
            communities[kv.first] = {kv.first};
            Node n;
            n.id = kv.first;
            n.nodes = {kv.first};
            node[kv.first] = n;
        }
    }
    Louvain() {}

    bool first_stage() {  // louvain.rs:72-117
        bool mod_inc = false;
        std::map<uint32_t, float> node_ids;
        std::vector<uint32_t> visit_ids;
        for (auto &kv : data) visit_ids.push_back(kv.first);
        std::sort(visit_ids.begin(), visit_ids.end());
        for (;;) {
            bool can_stop = true;
            for (uint32_t v_id : visit_ids) {
                uint32_t v_nid = node.at(v_id).id;
                node_ids.clear();
                const auto &nbrs = data.at(v_id);
                for (auto &kv : nbrs) {
                    uint32_t w_nid = node.at(kv.first).id;
                    if (node_ids.count(w_nid)) continue;
                    const auto &comm = communities.at(w_nid);
                    float sum = 0.f;
                    for (auto &kv2 : nbrs)
                        if (comm.count(kv2.first)) sum += kv2.second;
                    node_ids[w_nid] = sum;
                }
                // max_by(weight, then smaller id wins) louvain.rs:99-101
                bool have = false;
                uint32_t best_id = 0;
                float best_w = 0;
                for (auto &kv : node_ids) {
                    if (!have || kv.second > best_w || (kv.second == best_w && kv.first < best_id)) {
                        have = true;
                        best_id = kv.first;
                        best_w = kv.second;
                    }
                }
                if (have && best_w > 0.0f && best_id != v_nid) {
                    node.at(v_id).id = best_id;
                    communities.at(best_id).insert(v_id);
                    communities.at(v_nid).erase(v_id);
                    can_stop = false;
                    mod_inc = true;
                }
            }
            if (can_stop) break;
        }
        return mod_inc;
    }

    Louvain second_stage() {  // louvain.rs:119-195
        std::map<uint32_t, Node> new_node;
        std::map<uint32_t, std::set<uint32_t>> new_comm;
        std::vector<uint32_t> decluster_ids;
        for (auto &kv : communities) {
            if (kv.second.empty()) continue;
            uint32_t id = kv.first;
            const auto &nodes = kv.second;
            Node nn;
            nn.id = id;
            nn.weight = 0.f;
            for (uint32_t nid : nodes) {
                const Node &vertex = node.at(nid);
                nn.nodes.insert(vertex.nodes.begin(), vertex.nodes.end());
                nn.weight += vertex.weight;
                auto it = data.find(nid);
                if (it != data.end())
                    for (auto &e : it->second)
                        if (nodes.count(e.first)) nn.weight += e.second / 2.0f;
            }
            if (nn.weight < 0.f) {
                decluster_ids.push_back(id);
                g_order_exposure[1]++;
            }
            else {
                new_comm[id] = {id};
                new_node[id] = nn;
            }
        }
        for (uint32_t id : decluster_ids) {
            auto itc = communities.find(id);
            if (itc == communities.end()) fail("louvain: declustered community vanished (reference would panic)");
            std::set<uint32_t> nodes = itc->second;
            communities.erase(itc);
            for (uint32_t nid : nodes) {
                uint32_t new_nid = nid;
                while (new_comm.count(new_nid) || new_node.count(new_nid)) new_nid += 1;
                new_comm[new_nid] = {new_nid};
                Node nn;
                nn.id = new_nid;
                nn.weight = node.at(nid).weight;
                nn.nodes = node.at(nid).nodes;
                new_node[new_nid] = nn;
                communities[new_nid] = {nid};
            }
        }
        Graph nd;
        for (auto &c1 : communities) {
            if (c1.second.empty()) continue;
            for (auto &c2 : communities) {
                if (!(c2.first > c1.first) || c2.second.empty()) continue;
                float edge_weight = 0.0f;
                for (uint32_t vid : c1.second) {
                    auto it = data.find(vid);
                    if (it != data.end())
                        for (auto &e : it->second)
                            if (c2.second.count(e.first)) edge_weight += e.second;
                }
                if (edge_weight != 0.f) {
                    insert_data(nd, c1.first, c2.first, edge_weight);
                    insert_data(nd, c2.first, c1.first, edge_weight);
                }
            }
        }
        Louvain r;
        r.data = std::move(nd);
        r.communities = std::move(new_comm);
        r.node = std::move(new_node);
        return r;
    }

    void get_communities(Graph &out_data, std::vector<Node> &out) {  // louvain.rs:197-245
        for (auto &kv : communities) {
            if (kv.second.empty()) continue;
            float weight = 0.f;
            Node c;
            for (uint32_t vid : kv.second) {
                const Node &v = node.at(vid);
                c.nodes.insert(v.nodes.begin(), v.nodes.end());
                weight += v.weight;
                auto it = data.find(vid);
                if (it != data.end())
                    for (auto &e : it->second)
                        if (kv.second.count(e.first)) weight += e.second / 2.0f;
            }
            c.id = kv.first;
            c.weight = weight;
            out.push_back(std::move(c));
        }
        for (auto &c1 : out)
            for (auto &c2 : out) {
                if (!(c2.id > c1.id)) continue;
                float weight = 0.f;
                for (uint32_t n1 : communities.at(c1.id))
                    for (uint32_t n2 : communities.at(c2.id)) {
                        auto it = data.find(n1);
                        if (it != data.end()) {
                            auto jt = it->second.find(n2);
                            if (jt != it->second.end()) weight += jt->second;
                        }
                    }
                if (weight != 0.f) {
                    if (!(weight < 0.f)) fail("the weight of two conflicting community is not less than 0");
                    insert_data(out_data, c1.id, c2.id, weight);
                    insert_data(out_data, c2.id, c1.id, weight);
                }
            }
    }
};

// louvain.rs:290-356
std::vector<uint32_t> phase_communities(Graph data, const std::map<uint32_t, float> *ref_weight) {
    g_order_exposure[0]++;
    Louvain lv(std::move(data));
    for (;;) {  // louvain.rs:247-256
        bool mod_inc = lv.first_stage();
        if (mod_inc) lv = lv.second_stage();
        else break;
    }
    Graph cdata;
    std::vector<Node> communities;
    lv.get_communities(cdata, communities);
    std::map<uint32_t, std::pair<int64_t, float>> sort_key;  // community id -> key it was sorted by (exposure count)

    if (ref_weight) {
        struct Key {
            int32_t count;
            float weight;
        };
        std::vector<std::pair<Key, size_t>> keyed;
        for (size_t i = 0; i < communities.size(); i++) {
            int32_t count = 0;
            float weight = 0.f;
            for (uint32_t n : communities[i].nodes) {
                auto it = ref_weight->find(n);
                if (it != ref_weight->end()) {
                    if (it->second > 0.f) count += 1;
                    else if (it->second < 0.f) count -= 1;
                    weight += it->second;
                }
            }
            keyed.push_back({{count, weight}, i});
            sort_key[communities[i].id] = {count, weight};
        }
        std::stable_sort(keyed.begin(), keyed.end(), [](const std::pair<Key, size_t> &a, const std::pair<Key, size_t> &b) {
            // Reverse((count, weight)): descending
            if (a.first.count != b.first.count) return a.first.count > b.first.count;
            return a.first.weight > b.first.weight;
        });
        std::vector<Node> sorted;
        for (auto &k : keyed) sorted.push_back(std::move(communities[k.second]));
        communities.swap(sorted);
    } else {
        for (auto &c : communities) sort_key[c.id] = {0, c.weight};
        std::stable_sort(communities.begin(), communities.end(),
                         [](const Node &a, const Node &b) { return b.weight < a.weight; });
    }

    std::set<uint32_t> invalid_ids;
    for (size_t p = 0; p < communities.size(); p++) {
        if (invalid_ids.count(communities[p].id)) continue;
        auto it = cdata.find(communities[p].id);
        if (it != cdata.end()) {
            for (size_t q = p + 1; q < communities.size(); q++) {
                if (invalid_ids.count(communities[q].id)) continue;
                if (it->second.count(communities[q].id)) {
                    invalid_ids.insert(communities[q].id);
                    if (sort_key[communities[p].id] == sort_key[communities[q].id]) {  // order decided who is dropped
                        g_order_exposure[2]++;
                        g_order_exposure[3] += communities[p].nodes.size() + communities[q].nodes.size();
                    }
                }
            }
        }
    }
    std::vector<uint32_t> invalid_nodes;
    for (auto &c : communities)
        if (invalid_ids.count(c.id)) invalid_nodes.insert(invalid_nodes.end(), c.nodes.begin(), c.nodes.end());
    return invalid_nodes;
}

// main.rs:948-1015
// pair_weights (optional, stage dump): per read pair (a < b, a = 0 is the ref read) the number of heterozygous regions
// in which the two reads agree / differ, as #agree + #differ * (2^32 - 1) -- the pair loop of main.rs:953-992 before
// anything is derived from it
std::vector<uint32_t> phase_reads_by_lqseqs(const std::vector<LqSeqs> &lqseqs, bool asref, bool use_all_reads,
                                            std::map<uint64_t, int64_t> *pair_weights = nullptr) {
    Graph data, dif, ref_data;
    std::set<uint32_t> invalid_ids;
    for (auto &lqseq : lqseqs) {
        if (!lqseq.has_lable(LQSEQS_LABLE_HETE)) continue;
        for (size_t i = 0; i < lqseq.seqs.size(); i++) {
            const LqSeq &seq1 = lqseq.seqs[i];
            if (seq1.kscore == 0) continue;
            for (size_t j = i + 1; j < lqseq.seqs.size(); j++) {
                const LqSeq &seq2 = lqseq.seqs[j];
                if (seq2.kscore == 0) continue;
                float w = seq1.seq == seq2.seq ? 1.f : -1.f;
                if (pair_weights) {
                    const uint32_t a = std::min(seq1.order, seq2.order), b = std::max(seq1.order, seq2.order);
                    (*pair_weights)[(uint64_t)a << 32 | b] += w > 0.f ? 1 : ((int64_t)1 << 32) - 1;
                }
                if (seq1.order == 0) {
                    if (asref) insert_data(ref_data, seq1.order, seq2.order, w);
                    if (w < 0.f && !use_all_reads) invalid_ids.insert(seq2.order);
                    continue;
                }
                if (seq2.order == 0) fail("seq2 order is equal to 0");
                if (w == -1.f) {
                    insert_data(dif, seq1.order, seq2.order, -1.f);
                    insert_data(dif, seq2.order, seq1.order, -1.f);
                }
                insert_data(data, seq1.order, seq2.order, w);
                insert_data(data, seq2.order, seq1.order, w);
            }
        }
    }
    for (auto &n1 : dif)
        for (auto &n2 : n1.second)
            if (n2.second <= -3.f) assign_data(data, n1.first, n2.first, n2.second);
    if (!use_all_reads) {
        for (auto it = data.begin(); it != data.end();) {
            if (invalid_ids.count(it->first)) it = data.erase(it);
            else ++it;
        }
        for (auto &n1 : data)
            for (auto it = n1.second.begin(); it != n1.second.end();) {
                if (invalid_ids.count(it->first)) it = n1.second.erase(it);
                else ++it;
            }
    }
    const std::map<uint32_t, float> *rw = ref_data.empty() ? nullptr : &ref_data.begin()->second;
    std::vector<uint32_t> out = phase_communities(std::move(data), rw);
    out.insert(out.end(), invalid_ids.begin(), invalid_ids.end());
    return out;
}

/* --------------------------------------------------- consensus patching */

// main.rs:1017-1025 (usize wrap-around is the loop exit)
size_t get_lqseqs_next_idx_by_lable(const std::vector<LqSeqs> &lqseqs, size_t lqseqs_i, uint8_t lable) {
    lqseqs_i -= 1;
    while (lqseqs_i < lqseqs.size() && !lqseqs[lqseqs_i].has_lable(lable)) lqseqs_i -= 1;
    return lqseqs_i;
}

// main.rs:1027-1058
std::vector<ConsensusBase> update_consensus_with_lqseqs(const std::vector<LqSeqs> &lqseqs,
                                                        const std::vector<ConsensusBase> &consensus, uint8_t lable) {
    std::vector<ConsensusBase> out;
    out.reserve(consensus.size());
    size_t i = 0;
    size_t lqseqs_i = get_lqseqs_next_idx_by_lable(lqseqs, lqseqs.size(), lable);
    while (i < consensus.size()) {
        uint32_t p = consensus[i].pos;
        if (lqseqs_i < lqseqs.size() && p == lqseqs[lqseqs_i].start) {
            for (char base : lqseqs[lqseqs_i].sudoseed) out.push_back({p, (uint8_t)base});
            while (i < consensus.size() && consensus[i].pos <= lqseqs[lqseqs_i].end) i += 1;
            lqseqs_i = get_lqseqs_next_idx_by_lable(lqseqs, lqseqs_i, lable);
        } else {
            out.push_back(consensus[i]);
            i += 1;
        }
    }
    return out;
}

struct CnsView {
    const std::vector<ConsensusBase> &c;
    uint32_t pos(size_t i) const {
        if (i >= c.size()) fail("consensus index out of range in reupdate (reference would panic)");
        return c[i].pos;
    }
};

// main.rs:1068-1097 (not include s & e)
std::pair<size_t, size_t> iter_consensus_region(const std::vector<ConsensusBase> &consensus, size_t &idx, uint32_t s,
                                                uint32_t e) {
    CnsView v{consensus};
    size_t i = idx;
    while (v.pos(i) <= s) i += 1;
    while (v.pos(i) > s) i -= 1;
    i += 1;
    size_t si = i;
    while (v.pos(i) >= e) i -= 1;
    while (v.pos(i) < e) i += 1;
    i -= 1;
    idx = i;
    return {si, i + 1};
}

// main.rs:1100-1139 (not include p)
std::pair<size_t, size_t> iter_consensus_extend(const std::vector<ConsensusBase> &consensus, size_t &idx, uint32_t p,
                                                size_t l, bool toleft) {
    CnsView v{consensus};
    size_t si, ei;
    size_t i = idx;
    if (toleft) {
        while (v.pos(i) >= p) i -= 1;
        while (v.pos(i) < p) i += 1;
        idx = i;
        ei = i;
        si = i > l ? i - l : 0;
    } else {
        while (v.pos(i) <= p) i += 1;
        while (v.pos(i) > p) i -= 1;
        idx = i;
        si = i + 1;
        ei = i + l < consensus.size() ? i + l + 1 : consensus.size();
    }
    return {si, ei};
}

// main.rs:1060-1420
std::vector<ConsensusBase> reupdate_consensus_with_lqseqs(std::vector<LqSeqs> &lqseqs,
                                                          const std::vector<ConsensusBase> &consensus,
                                                          KmerInfo &kmer_info, uint16_t min_kmer_count,
                                                          size_t iter_count) {
    kmer_info.clear();
    const size_t ksize = kmer_info.ksize;
    std::vector<size_t> rech_idxs;
    for (size_t i = lqseqs.size(); i-- > 0;)
        if (lqseqs[i].has_lable(LQSEQS_LABLE_RECH)) rech_idxs.push_back(i);

    std::string buf;
    auto append_cns = [&](size_t si, size_t ei) {
        for (size_t x = si; x < ei; x++) buf.push_back((char)consensus[x].base);
    };
    // main.rs:1141-1176
    auto chain = [&](const std::vector<size_t> &choice, size_t sj, size_t &idx, size_t si_l, size_t ei_l, size_t si_r,
                     size_t ei_r) {
        buf.clear();
        append_cns(si_l, ei_l);
        for (size_t i = 0; i < choice.size(); i++) {
            const std::string &seq = lqseqs[rech_idxs[sj + i]].seqs[choice[i]].seq;
            if (i < choice.size() - 1) {
                uint32_t s = lqseqs[rech_idxs[sj + i]].end;
                uint32_t e = lqseqs[rech_idxs[sj + i + 1]].start;
                buf += seq;
                if (s + 1 != e) {
                    auto r = iter_consensus_region(consensus, idx, s, e);
                    append_cns(r.first, r.second);
                }
            } else {
                buf += seq;
                append_cns(si_r, ei_r);
            }
        }
    };
    // itertools multi_cartesian_product (main.rs:1244,1327): lexicographic, last iterator fastest
    auto for_each_product = [&](size_t sj, size_t ej, const std::function<void(const std::vector<size_t> &)> &f) {
        size_t n = ej - sj;
        std::vector<size_t> lens(n), choice(n, 0);
        for (size_t x = 0; x < n; x++) {
            lens[x] = lqseqs[rech_idxs[sj + x]].seqs.size();
            if (lens[x] == 0) return;
        }
        for (;;) {
            f(choice);
            size_t d = n;
            while (d-- > 0) {
                if (++choice[d] < lens[d]) break;
                choice[d] = 0;
            }
            if (d == (size_t)-1) break;
        }
    };
    auto min_score = [&](const std::string &s) -> uint16_t {
        bool any = false;
        uint16_t mn = 0;
        iter2kmer((const uint8_t *)s.data(), (const uint8_t *)s.data() + s.size(), ksize, [&](uint64_t x) {
            int g = kmer_info.get(kmer_info.to_hash(x));
            uint16_t v = g < 0 ? 0 : (uint16_t)g;
            if (!any || v < mn) mn = v;
            any = true;
        });
        return any ? mn : 0;
    };

    for (int pass = 0; pass < 2; pass++) {
        if (pass == 1) kmer_info.retrieve_kmers(min_kmer_count);  // main.rs:1267
        size_t idx = 0, sj = 0, ej;
        std::vector<std::tuple<size_t, size_t, uint16_t>> kscore_buf;
        while (sj < rech_idxs.size()) {
            ej = sj + 1;
            while (ej < rech_idxs.size() &&
                   lqseqs[rech_idxs[ej]].start < lqseqs[rech_idxs[ej - 1]].end + kmer_info.ksize) {
                ej += 1;
                if (ej > sj + 5) break;
            }
            auto L = iter_consensus_extend(consensus, idx, lqseqs[rech_idxs[sj]].start, ksize - 1, true);
            auto R = iter_consensus_extend(consensus, idx, lqseqs[rech_idxs[ej - 1]].end, ksize - 1, false);
            if (ej == sj + 1) {
                for (auto &seq : lqseqs[rech_idxs[sj]].seqs) {
                    buf.clear();
                    append_cns(L.first, L.second);
                    buf += seq.seq;
                    append_cns(R.first, R.second);
                    if (pass == 0) {
                        iter2kmer((const uint8_t *)buf.data(), (const uint8_t *)buf.data() + buf.size(), ksize,
                                  [&](uint64_t kmer) { kmer_info.insert(kmer_info.to_hash(kmer)); });
                    } else {
                        seq.kscore = min_score(buf);
                    }
                }
            } else {
                kscore_buf.clear();
                for_each_product(sj, ej, [&](const std::vector<size_t> &choice) {
                    chain(choice, sj, idx, L.first, L.second, R.first, R.second);
                    if (pass == 0) {
                        iter2kmer((const uint8_t *)buf.data(), (const uint8_t *)buf.data() + buf.size(), ksize,
                                  [&](uint64_t kmer) { kmer_info.insert(kmer_info.to_hash(kmer)); });
                    } else {
                        uint16_t kscore = min_score(buf);
                        if (kscore > 0)
                            for (size_t i = 0; i < choice.size(); i++)
                                kscore_buf.emplace_back(rech_idxs[sj + i], choice[i], kscore);
                    }
                });
                if (pass == 1) {
                    for (size_t x = sj; x < ej; x++)
                        for (auto &seq : lqseqs[rech_idxs[x]].seqs) seq.kscore = 0;
                    for (auto &t : kscore_buf) lqseqs[std::get<0>(t)].seqs[std::get<1>(t)].kscore = std::get<2>(t);
                }
            }
            sj = ej;
        }
    }

    // main.rs:1371-1406
    for (auto &lqseq : lqseqs) {
        if (!lqseq.has_lable(LQSEQS_LABLE_RECH)) continue;
        size_t c = 0, valid_count = 0;
        for (size_t p = 0; p < lqseq.seqs.size(); p++) {
            const LqSeq &seq = lqseq.seqs[p];
            if (seq.kscore != 0) {
                if (c == 0 || seq.order == 0) c = p + 1;
                valid_count += 1;
            }
        }
        if (valid_count > 1) lqseq.set_lable(LQSEQS_LABLE_TEMP);
        if (c != 0) {
            lqseq.sudoseed = lqseq.seqs[c - 1].seq;
        } else if (iter_count == 1) {
            size_t i = 0;
            for (size_t p = 0; p < lqseq.seqs.size(); p++)
                if (lqseq.seqs[p].order == 0) {
                    i = p;
                    break;
                }
            if (lqseq.seqs.empty()) fail("RECH region without candidates (reference would panic)");
            lqseq.sudoseed = lqseq.seqs[i].seq;
        }
    }
    std::vector<ConsensusBase> out = update_consensus_with_lqseqs(lqseqs, consensus, LQSEQS_LABLE_RECH);
    for (auto &lqseq : lqseqs) {  // main.rs:1411-1417
        if (!lqseq.has_lable(LQSEQS_LABLE_RECH)) continue;
        if (lqseq.has_lable(LQSEQS_LABLE_TEMP)) lqseq.unset_lable(LQSEQS_LABLE_TEMP);
        else lqseq.unset_lable(LQSEQS_LABLE_RECH);
    }
    return out;
}

}  // namespace

/* ---------------------------------------------------------------- job */

struct np2o_job {
    std::vector<uint8_t> tseq;
    std::vector<uint8_t> bam;
    np2o_opts opt;
    std::vector<const np2o_table *> tables;
    int32_t dump_iter = 0;
    double seconds = 0;

    // dumps
    std::vector<int32_t> d_rec_idx;
    std::vector<uint32_t> d_ts, d_te;
    std::vector<uint64_t> d_nib_off;
    std::vector<uint8_t> d_nib, d_blank;
    std::vector<uint64_t> d_msa_off;
    std::vector<uint16_t> d_msa_bases, d_msa_delta;
    std::vector<uint32_t> d_msa_count, d_msa_besti;
    std::vector<uint32_t> d_dp_pos;
    std::vector<uint8_t> d_dp_base, d_dp_flags;
    std::vector<uint32_t> d_reg_start, d_reg_end;
    std::vector<uint8_t> d_reg_lable;
    std::vector<uint64_t> d_can_roff, d_can_kmer, d_can_seq_off;
    std::vector<uint32_t> d_can_order;
    std::vector<uint16_t> d_can_kscore;
    std::vector<uint8_t> d_can_seq;
    std::vector<uint32_t> d_dropped;
    std::vector<uint64_t> d_pair_key;
    std::vector<int64_t> d_pair_val;
    std::vector<uint32_t> d_cns_pos;
    std::vector<uint8_t> d_cns_base;

    void run();
    std::vector<ConsensusBase> get_cns_from_align_tags(std::vector<Msa> &msas, std::vector<AlignSeq> &alignseqs,
                                                       std::vector<KmerInfo> &yak, bool out_cns, bool dump,
                                                       bool &have);
};

// main.rs:1645-1687 + 1555-1643 + 1422-1553
std::vector<ConsensusBase> np2o_job::get_cns_from_align_tags(std::vector<Msa> &msas, std::vector<AlignSeq> &alignseqs,
                                                             std::vector<KmerInfo> &yak, bool out_cns, bool dump,
                                                             bool &have) {
    static const Kmer default_kmer;  // main.rs:1651
    const Kmer *global_best_kmer = &default_kmer;
    const size_t L = msas.size();

    for (size_t p = 0; p < L; p++) {  // main.rs:1653-1684
        const Msa &msa = msas[p];
        for (const Kmer &kmer : msa.kmers) {
            AlignBase base1, base2, base3;
            kmer.get_bases((uint32_t)p, base1, base2, base3);
            int64_t coverage = msa.coverage();
            uint32_t besti = 0;
            int64_t kmer_score;
            if (base2.is_head()) {
                kmer_score = 10 * (int64_t)kmer.count - 4 * coverage;
            } else {
                kmer_score = INT64_MIN >> 1;
                if (base2.t_pos >= L) fail("3-mer predecessor position out of range");
                const Msa &pm = msas[base2.t_pos];
                // Msa::get main.rs:209-225
                uint8_t base23 = (uint8_t)(base1.q_base << 4 | base2.q_base);
                uint16_t delta23 = base1.t_pos == base2.t_pos ? 1 : 0;
                for (size_t pi = 0; pi < pm.kmers.size(); pi++) {
                    const Kmer &v = pm.kmers[pi];
                    if (!((uint8_t)v.bases == base23 && (v.bases >> 12 & 1) == delta23)) continue;
                    AlignBase pb1, pb2, pb3;
                    v.get_bases(base2.t_pos, pb1, pb2, pb3);
                    if (!(pb2 == base1 && pb3 == base2)) continue;
                    if (base2.t_pos >= 3 && pb1.is_head()) continue;  // main.rs:1666-1668
                    int64_t score = v.score + 10 * (int64_t)kmer.count - 4 * coverage;
                    if (score > kmer_score || (score == kmer_score && pb1.q_base != 4)) {
                        kmer_score = score;
                        besti = (uint32_t)pi;
                    }
                }
            }
            kmer.score = kmer_score;
            kmer.besti = besti;
            if (p == L - 1 && kmer_score >= global_best_kmer->score) global_best_kmer = &kmer;
        }
    }

    if (dump) {
        d_msa_off.assign(L + 1, 0);
        for (size_t p = 0; p < L; p++) {
            d_msa_off[p + 1] = d_msa_off[p] + msas[p].kmers.size();
            for (auto &k : msas[p].kmers) {
                d_msa_bases.push_back(k.bases);
                d_msa_delta.push_back(k.delta);
                d_msa_count.push_back(k.count);
                d_msa_besti.push_back(k.besti);
            }
        }
    }

    // generate_cns_from_best_score_lq main.rs:1555-1643
    std::vector<LqSeqs> lqseqs;
    std::vector<ConsensusBase> consensusbases;
    std::vector<uint8_t> cflags;
    consensusbases.reserve(L);
    const int64_t hq_min_qv = 95;
    const size_t lq_min_length = 2;
    bool has_lq = false;
    size_t lq_s = SIZE_MAX, lq_e = 0, p = 0;
    AlignBase base1, base2, base3;
    global_best_kmer->get_bases((uint32_t)L - 1, base1, base2, base3);
    for (;;) {
        if (base3.q_base != 4) {
            if (base3.t_pos >= L) fail("backtrack left the contig");
            int64_t coverage = msas[base3.t_pos].coverage();
            if (coverage == 0) fail("zero coverage in qv (reference would panic: division by zero)");
            int64_t qv = (int64_t)global_best_kmer->count * 100 / coverage;
            consensusbases.push_back({base3.t_pos, SEQ_NUM[base3.q_base]});
            cflags.push_back((uint8_t)((qv < hq_min_qv ? 1 : 0) | (coverage < 2 ? 2 : 0)));
            if (coverage < 2) {
                has_lq = false;
                lq_s = SIZE_MAX;
            } else if (qv < hq_min_qv) {
                if (lq_s == SIZE_MAX) lq_s = p;
                lq_e = p;
                has_lq = true;
            } else if (has_lq && p - lq_e > 2 * lq_min_length &&
                       consensusbases[p - 1].pos != consensusbases[p - 2].pos &&
                       consensusbases[p - 1].base != consensusbases[p - 2].base) {
                lq_e = p - 2;
                lq_s = lq_s > lq_min_length ? lq_s - lq_min_length : 1;
                while (lq_s > 1 && (consensusbases[lq_s - 1].pos == consensusbases[lq_s].pos ||
                                    consensusbases[lq_s - 1].base == consensusbases[lq_s].base))
                    lq_s -= 1;
                size_t lqseqs_index = lqseqs.size();
                if (lqseqs_index >= 1 && consensusbases[lq_s].pos >= lqseqs[lqseqs_index - 1].start) {
                    lqseqs[lqseqs_index - 1].start = consensusbases[lq_e].pos;
                } else {
                    LqSeqs r;
                    r.end = consensusbases[lq_s].pos;
                    r.start = consensusbases[lq_e].pos;
                    lqseqs.push_back(std::move(r));
                }
                has_lq = false;
                lq_s = SIZE_MAX;
            }
            p += 1;
        }
        if (base2.is_head()) break;
        if (base2.t_pos >= L) fail("backtrack left the contig");
        const Msa &pm = msas[base2.t_pos];
        if (global_best_kmer->besti >= pm.kmers.size()) fail("besti out of range (reference would panic)");
        global_best_kmer = &pm.kmers[global_best_kmer->besti];
        AlignBase nb1;
        global_best_kmer->get_bases(base2.t_pos, nb1, base2, base3);
    }
    std::reverse(consensusbases.begin(), consensusbases.end());
    std::reverse(cflags.begin(), cflags.end());

    if (dump) {
        for (size_t i = 0; i < consensusbases.size(); i++) {
            d_dp_pos.push_back(consensusbases[i].pos);
            d_dp_base.push_back(consensusbases[i].base);
            d_dp_flags.push_back(cflags[i]);
        }
        for (auto &r : lqseqs) {
            d_reg_start.push_back(r.start);
            d_reg_end.push_back(r.end);
        }
    }

    have = true;
    if (lqseqs.empty()) return consensusbases;  // main.rs:1638-1640

    // generate_lqseqs_from_tags_kmer main.rs:1422-1553
    std::vector<AlignBase> align_bases;
    KmerInfo &kmer_info = yak[0];
    const uint64_t ksize = kmer_info.ksize;
    if (ksize >= 32) fail("the smallest yak table must have k < 32 (main.rs:1432-1434)");
    const uint64_t shift = 2 * (ksize - 1);
    const uint64_t mask = (1ULL << (2 * ksize)) - 1;
    uint64_t kmers[2] = {0, 0};
    uint64_t l;
    size_t j, s = lqseqs.size() - 1;
    for (size_t idx = 0; idx < alignseqs.size(); idx++) {
        const AlignSeq &ab = alignseqs[idx];
        if (ab.align_bases.empty()) continue;
        while (s > 0 && lqseqs[s].start < ab.aln_t_s) s -= 1;
        if (lqseqs[s].start < ab.aln_t_s || lqseqs[s].end > ab.aln_t_e) continue;
        j = s;
        while (j > 0 && lqseqs[j].end <= ab.aln_t_e) j -= 1;
        if (lqseqs[j].end > ab.aln_t_e) j += 1;

        align_bases.clear();
        size_t pp = 0;
        AlignBase align_base;
        while (ab.get_align_tag(pp, align_base)) {
            align_bases.push_back(align_base);
            if (align_base.t_pos > lqseqs[j].end + (uint32_t)ksize) break;
        }
        for (size_t ri = j; ri <= s; ri++) {
            LqSeqs &lqseq = lqseqs[ri];
            if (lqseq.seqs.size() >= LQSEQ_MAX_CAN_COUNT) continue;
            l = 0;
            kmers[0] = kmers[1] = 0;
            std::string seq;
            size_t from = (size_t)lqseq.start - (size_t)ab.aln_t_s;
            if (from > align_bases.size()) fail("slice start beyond decoded columns (reference would panic)");
            for (size_t x = from; x < align_bases.size(); x++) {
                const AlignBase &a = align_bases[x];
                if (a.t_pos >= lqseq.start && a.q_base != 4) {
                    if (a.t_pos <= lqseq.end) seq.push_back((char)SEQ_NUM[a.q_base]);
                    if (l < ksize) {
                        kmers[0] = (kmers[0] << 2 | (uint64_t)a.q_base) & mask;
                        kmers[1] = (kmers[1] >> 2) | (3 ^ (uint64_t)a.q_base) << shift;
                        l += 1;
                    }
                    if (a.t_pos > lqseq.end && l >= ksize) break;
                }
            }
            uint64_t kmer = l >= ksize ? (kmers[0] < kmers[1] ? kmers[0] : kmers[1]) : INVALID_KMER;
            if (!seq.empty()) {
                LqSeq q;
                q.order = (uint32_t)idx;
                q.kscore = 0;
                q.kmer = kmer != INVALID_KMER ? kmer_info.to_hash(kmer) : INVALID_KMER;
                q.seq = std::move(seq);
                lqseq.seqs.push_back(std::move(q));
            }
        }
    }

    retrieve_kmer_count(lqseqs, kmer_info, (uint16_t)opt.min_kmer_count);  // main.rs:1525

    if (dump) {
        d_can_roff.push_back(0);
        d_can_seq_off.push_back(0);
        for (auto &r : lqseqs) {
            for (auto &c : r.seqs) {
                d_can_order.push_back(c.order);
                d_can_kscore.push_back(c.kscore);
                d_can_kmer.push_back(c.kmer);
                d_can_seq.insert(d_can_seq.end(), c.seq.begin(), c.seq.end());
                d_can_seq_off.push_back(d_can_seq.size());
            }
            d_can_roff.push_back(d_can_order.size());
        }
    }

    if (out_cns) {  // main.rs:1527-1543
        fill_seed_lqseqs(lqseqs, opt.max_indel_len);
        std::vector<ConsensusBase> consensus = update_consensus_with_lqseqs(lqseqs, consensusbases, LQSEQS_LABLE_SUCC);
        for (size_t pi = 0; pi < yak.size(); pi++)
            consensus = reupdate_consensus_with_lqseqs(lqseqs, consensus, yak[pi], (uint16_t)opt.min_kmer_count, pi + 1);
        if (dump)
            for (auto &r : lqseqs) d_reg_lable.push_back(r.lable);
        return consensus;
    } else {  // main.rs:1544-1552
        mark_hete_lqseqs(lqseqs);
        if (dump)
            for (auto &r : lqseqs) d_reg_lable.push_back(r.lable);
        std::map<uint64_t, int64_t> pw;
        std::vector<uint32_t> invalid_ids = phase_reads_by_lqseqs(lqseqs, opt.model == 0, opt.use_all_reads != 0, dump ? &pw : nullptr);
        if (dump)
            for (auto &kv : pw) {
                d_pair_key.push_back(kv.first);
                d_pair_val.push_back(kv.second);
            }
        std::sort(invalid_ids.begin(), invalid_ids.end());
        invalid_ids.erase(std::unique(invalid_ids.begin(), invalid_ids.end()), invalid_ids.end());
        for (uint32_t id : invalid_ids) {
            alignseqs.at(id).align_bases.clear();
            d_dropped.push_back(id);
        }
        have = false;
        return {};
    }
}

// worker closure main.rs:1726-1838
void np2o_job::run() {
    const size_t tlen = tseq.size();
    // -S: secondary records are expected to carry their SEQ already (the caller recovered it from the primary
    // record, secondary.rs:85-150 + main.rs:1775-1783); the filter below then keeps them (main.rs:1764).
    if (tables.empty()) fail("Missing yak file!");
    if (opt.iter_count == 0) fail("iter_count must be >= 1");
    if (tlen < opt.min_ctg_len) {  // main.rs:1727-1730
        for (size_t p = 0; p < tlen; p++) {
            d_cns_pos.push_back((uint32_t)p);
            d_cns_base.push_back(tseq[p]);
        }
        return;
    }
    std::vector<KmerInfo> yak;
    for (auto *t : tables) yak.emplace_back(t);
    std::stable_sort(yak.begin(), yak.end(), [](const KmerInfo &a, const KmerInfo &b) { return a.ksize < b.ksize; });  // option.rs:238

    Alignment aln;
    std::vector<AlignSeq> alignseqs;
    std::vector<Msa> msas(tlen);
    aln.aln_t_e = (uint32_t)tlen;
    aln.aln_q_e = (uint32_t)tlen;
    aln.q_aln_str.assign((const char *)tseq.data(), tlen);
    aln.t_aln_str.assign((const char *)tseq.data(), tlen);
    alignseqs.push_back(AlignSeq::make(aln));
    std::vector<int32_t> rec_of;
    rec_of.push_back(-1);

    // BAM records main.rs:1749-1816
    int64_t pre_tid = 0, pre_pos = 0;
    uint64_t off = 0;
    int32_t rec = -1;
    while (off + 4 <= bam.size()) {
        rec++;
        int32_t block_size;
        memcpy(&block_size, &bam[off], 4);
        if (block_size < 32 || off + 4 + (uint64_t)block_size > bam.size()) fail("BAM/SAM parsing failed!");
        const uint8_t *r = &bam[off + 4];
        off += 4 + (uint64_t)block_size;
        int32_t refID, pos, l_seq;
        uint8_t l_read_name, mapq;
        uint16_t n_cigar, flag;
        memcpy(&refID, r, 4);
        memcpy(&pos, r + 4, 4);
        l_read_name = r[8];
        mapq = r[9];
        memcpy(&n_cigar, r + 12, 2);
        memcpy(&flag, r + 14, 2);
        memcpy(&l_seq, r + 16, 4);
        if (32 + (uint64_t)l_read_name + 4ull * n_cigar + ((uint64_t)l_seq + 1) / 2 > (uint64_t)block_size)
            fail("BAM/SAM parsing failed!");
        std::vector<uint32_t> cigar(n_cigar);
        memcpy(cigar.data(), r + 32 + l_read_name, 4ull * n_cigar);
        const uint8_t *seq4 = r + 32 + l_read_name + 4ull * n_cigar;

        if (!(refID > pre_tid || (int64_t)pos >= pre_pos)) fail("Unsorted input file!");  // main.rs:1753-1756

        // rust-htslib accessors
        uint64_t rlen = 0, rspan = 0;
        for (uint32_t c : cigar) {
            uint32_t l = c >> 4, op = c & 15;
            if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8 || op == 5) rlen += l;
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rspan += l;
        }
        int64_t ref_start = pos;
        int64_t ref_end = (flag & 4) || n_cigar == 0 || rspan == 0 ? ref_start + 1 : ref_start + (int64_t)rspan;
        bool is_secondary = flag & 0x100, is_supp = flag & 0x800;
        if ((flag & 0x404) != 0 || (int16_t)mapq <= (int16_t)opt.min_map_qual || rlen <= opt.min_read_len ||
            (is_secondary && !opt.use_secondary) || (is_supp && !opt.use_supplementary) ||
            (ref_end - ref_start <
             std::max((int64_t)opt.min_map_len, (int64_t)((float)rlen * opt.min_map_fra)))) {  // main.rs:1759-1771
            continue;
        }
        aln.clear();
        if (pos < 0 || (uint64_t)pos > tlen) fail("alignment starts outside the contig");
        aln.aln_t_s = (uint32_t)pos;
        aln.fill_with_cigar(cigar.data(), n_cigar, tseq.data() + pos, tlen - pos, seq4, (uint32_t)l_seq);
        bool is_clip = (uint32_t)(aln.aln_q_e - aln.aln_q_s + opt.max_clip_len) < (uint32_t)rlen;  // main.rs:1796-1797
        aln.trim(8);
        if (aln.aln_len() <= opt.min_map_len) continue;  // main.rs:1800
        AlignSeq alignseq = AlignSeq::make(aln);
        if (is_clip) {
            if (tlen < 500000) continue;  // main.rs:1807-1810
            alignseq.set_lable();
        }
        alignseqs.push_back(std::move(alignseq));
        rec_of.push_back(rec);
        pre_tid = refID;
        pre_pos = ref_start;
    }
    if (off != bam.size()) fail("BAM/SAM parsing failed!");
    filter_alignseqs_by_clip(alignseqs);  // main.rs:1817

    // dump reads
    d_nib_off.push_back(0);
    for (size_t i = 0; i < alignseqs.size(); i++) {
        d_rec_idx.push_back(rec_of[i]);
        d_ts.push_back(alignseqs[i].aln_t_s);
        d_te.push_back(alignseqs[i].aln_t_e);
        d_blank.push_back(alignseqs[i].align_bases.empty() ? 1 : 0);
        if (i > 0) d_nib.insert(d_nib.end(), alignseqs[i].align_bases.begin(), alignseqs[i].align_bases.end());
        d_nib_off.push_back(d_nib.size());
    }

    // main.rs:1819-1836
    uint32_t i = 0;
    for (;;) {
        update_msas(msas, alignseqs);
        for (auto &m : msas) m.sort();
        bool have = false;
        bool dump = (int32_t)i == dump_iter;
        if (i + 1 == opt.iter_count) {
            std::vector<ConsensusBase> cns = get_cns_from_align_tags(msas, alignseqs, yak, true, dump, have);
            for (auto &b : cns) {
                d_cns_pos.push_back(b.pos);
                d_cns_base.push_back(b.base);
            }
            break;
        } else {
            get_cns_from_align_tags(msas, alignseqs, yak, false, dump, have);
            for (auto &m : msas) m.kmers.clear();
        }
        i += 1;
    }
}

/* ---------------------------------------------------------------- C ABI */

extern "C" {

const char *np2o_last_error(void) { return g_err.c_str(); }

uint64_t np2o_yak_hash64(uint64_t key, uint64_t mask) { return yak_hash64(key, mask); }
uint64_t np2o_yak_hash64_64(uint64_t key) { return yak_hash64_64(key); }
uint64_t np2o_yak_hash_long(const uint64_t x[4]) { return yak_hash_long(x); }

int64_t np2o_seq_hashes(const uint8_t *seq, uint64_t len, uint32_t k, uint64_t *out, uint64_t cap) {
    try {
        if (k == 0 || k > 63) fail("k out of range");
        uint64_t n = 0;
        const uint64_t kmask = k < 32 ? (1ULL << (2 * (uint64_t)k)) - 1 : 0;
        iter2kmer(seq, seq + len, k, [&](uint64_t x) {
            uint64_t h = k < 32 ? yak_hash64(x, kmask) : x;
            if (n < cap) out[n] = h;
            n++;
        });
        return (int64_t)n;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

np2o_table *np2o_table_load(const char *path) {
    FILE *fp = fopen(path, "rb");
    if (!fp) {
        g_err = std::string("cannot open ") + path;
        return nullptr;
    }
    np2o_table *t = nullptr;
    try {
        uint8_t hdr[16];
        if (fread(hdr, 1, 16, fp) != 16 || memcmp(hdr, "YAK\2", 4) != 0)
            fail("The input binary k-mer dump file is incompatible.");  // kmer.rs:76-80
        uint32_t k, pre, cb;
        memcpy(&k, hdr + 4, 4);
        memcpy(&pre, hdr + 8, 4);
        memcpy(&cb, hdr + 12, 4);
        if (cb != 10) fail("different YAK_COUNTER_BITS");  // kmer.rs:90
        if (pre != 10) fail("yak prefix bits must be 10 (SURVEY §8b implicit contract)");
        fseek(fp, 0, SEEK_END);
        long fsz = ftell(fp);
        fseek(fp, 16, SEEK_SET);
        t = new np2o_table();
        t->k = k;
        t->pre = pre;
        t->reserve((uint64_t)fsz / 8);
        std::vector<uint64_t> buf;
        for (uint32_t b = 0; b < (1u << pre); b++) {
            uint32_t cs[2];
            if (fread(cs, 4, 2, fp) != 2) fail("Failed to parse the dump file");
            buf.resize(cs[1]);
            if (cs[1] && fread(buf.data(), 8, cs[1], fp) != cs[1]) fail("Failed to parse the dump file");
            for (uint64_t key : buf) t->put((key >> 10) << pre | b, (uint16_t)(key & 1023));
        }
        fclose(fp);
        return t;
    } catch (const std::exception &e) {
        g_err = e.what();
        fclose(fp);
        delete t;
        return nullptr;
    }
}

np2o_table *np2o_table_from_arrays(uint32_t k, const uint64_t *hashes, const uint16_t *counts, uint64_t n) {
    np2o_table *t = new np2o_table();
    t->k = k;
    t->reserve(n);
    for (uint64_t i = 0; i < n; i++) t->put(hashes[i], counts[i] & 1023);
    return t;
}

void np2o_table_destroy(np2o_table *t) { delete t; }
uint32_t np2o_table_k(const np2o_table *t) { return t->k; }
uint64_t np2o_table_size(const np2o_table *t) { return t->n; }
void np2o_table_set_stream_scan(np2o_table *t, int on) { t->stream_scan = on != 0; }

void np2o_table_lookup(const np2o_table *t, const uint64_t *hashes, uint64_t n, uint32_t min_count, uint16_t *out) {
    // insert + retrieve_kmers + get, batch form (kmer.rs:113-170)
    KmerInfo ki(t);
    for (uint64_t i = 0; i < n; i++) ki.insert(hashes[i]);
    ki.retrieve_kmers((uint16_t)min_count);
    for (uint64_t i = 0; i < n; i++) {
        int g = ki.get(hashes[i]);
        out[i] = g < 0 ? 0 : (uint16_t)g;
    }
}

np2o_job *np2o_job_create(const uint8_t *tseq, uint32_t tlen, const uint8_t *bam, uint64_t bam_len,
                          const np2o_opts *opts, np2o_table *const *tables, uint32_t n_tables) {
    np2o_job *j = new np2o_job();
    j->tseq.assign(tseq, tseq + tlen);
    j->bam.assign(bam, bam + bam_len);
    j->opt = *opts;
    for (uint32_t i = 0; i < n_tables; i++) j->tables.push_back(tables[i]);
    return j;
}

int np2o_job_run(np2o_job *job, int32_t dump_iter) {
    job->dump_iter = dump_iter;
    auto t0 = std::chrono::steady_clock::now();
    try {
        job->run();
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
    job->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return 0;
}

void np2o_job_destroy(np2o_job *job) { delete job; }

uint64_t np2o_get_reads(np2o_job *j, const int32_t **rec_idx, const uint32_t **t_s, const uint32_t **t_e,
                        const uint64_t **nib_off, const uint8_t **nib, const uint8_t **blank) {
    *rec_idx = j->d_rec_idx.data();
    *t_s = j->d_ts.data();
    *t_e = j->d_te.data();
    *nib_off = j->d_nib_off.data();
    *nib = j->d_nib.data();
    *blank = j->d_blank.data();
    return j->d_rec_idx.size();
}
uint64_t np2o_get_msa(np2o_job *j, const uint64_t **off, const uint16_t **bases, const uint16_t **delta,
                      const uint32_t **count, const uint32_t **besti) {
    *off = j->d_msa_off.data();
    *bases = j->d_msa_bases.data();
    *delta = j->d_msa_delta.data();
    *count = j->d_msa_count.data();
    *besti = j->d_msa_besti.data();
    return j->d_msa_bases.size();
}
uint64_t np2o_get_dp_consensus(np2o_job *j, const uint32_t **pos, const uint8_t **base, const uint8_t **flags) {
    *pos = j->d_dp_pos.data();
    *base = j->d_dp_base.data();
    *flags = j->d_dp_flags.data();
    return j->d_dp_pos.size();
}
uint64_t np2o_get_regions(np2o_job *j, const uint32_t **start, const uint32_t **end, const uint8_t **lable) {
    *start = j->d_reg_start.data();
    *end = j->d_reg_end.data();
    *lable = j->d_reg_lable.data();
    return j->d_reg_start.size();
}
uint64_t np2o_get_candidates(np2o_job *j, const uint64_t **roff, const uint32_t **order, const uint16_t **kscore,
                             const uint64_t **kmer, const uint64_t **seq_off, const uint8_t **seq) {
    *roff = j->d_can_roff.data();
    *order = j->d_can_order.data();
    *kscore = j->d_can_kscore.data();
    *kmer = j->d_can_kmer.data();
    *seq_off = j->d_can_seq_off.data();
    *seq = j->d_can_seq.data();
    return j->d_can_order.size();
}
uint64_t np2o_get_dropped(np2o_job *j, const uint32_t **ids) {
    *ids = j->d_dropped.data();
    return j->d_dropped.size();
}
uint64_t np2o_get_pair_weights(np2o_job *j, const uint64_t **keys, const int64_t **vals) {
    *keys = j->d_pair_key.data();
    *vals = j->d_pair_val.data();
    return j->d_pair_key.size();
}
/* Test seam: the tail of phase_reads_by_lqseqs (main.rs:994-1015) + Louvain + phase_communities on pair weights that
 * were already summed per read pair.  keys[e] = a << 32 | b (a < b), vals[e] = #agree + #differ * ((1 << 32) - 1).
 * Feeding the sums through insert_data / assign_data gives the same maps as the per-region +-1 updates. */
int64_t np2o_debug_phase(const uint64_t *keys, const int64_t *vals, uint64_t n, uint32_t model, uint32_t use_all_reads,
                         uint32_t *out, uint64_t cap) {
    try {
        const bool asref = model == 0, use_all = use_all_reads != 0;
        Graph data, dif, ref_data;
        std::set<uint32_t> invalid_ids;
        for (uint64_t e = 0; e < n; e++) {
            const uint32_t a = (uint32_t)(keys[e] >> 32), b = (uint32_t)keys[e];
            const int64_t ndif = (vals[e] + (1LL << 31)) >> 32;
            const int64_t sum = vals[e] - (ndif << 32);
            if (a == 0) {
                if (asref) insert_data(ref_data, a, b, (float)sum);
                if (ndif > 0 && !use_all) invalid_ids.insert(b);
                continue;
            }
            if (ndif > 0) {
                insert_data(dif, a, b, -(float)ndif);
                insert_data(dif, b, a, -(float)ndif);
            }
            insert_data(data, a, b, (float)sum);
            insert_data(data, b, a, (float)sum);
        }
        for (auto &n1 : dif)
            for (auto &n2 : n1.second)
                if (n2.second <= -3.f) assign_data(data, n1.first, n2.first, n2.second);
        if (!use_all) {
            for (auto it = data.begin(); it != data.end();) {
                if (invalid_ids.count(it->first)) it = data.erase(it);
                else ++it;
            }
            for (auto &n1 : data)
                for (auto it = n1.second.begin(); it != n1.second.end();) {
                    if (invalid_ids.count(it->first)) it = n1.second.erase(it);
                    else ++it;
                }
        }
        const std::map<uint32_t, float> *rw = ref_data.empty() ? nullptr : &ref_data.begin()->second;
        std::vector<uint32_t> res = phase_communities(std::move(data), rw);
        res.insert(res.end(), invalid_ids.begin(), invalid_ids.end());
        std::sort(res.begin(), res.end());
        res.erase(std::unique(res.begin(), res.end()), res.end());
        for (uint64_t i = 0; i < res.size() && i < cap; i++) out[i] = res[i];
        return (int64_t)res.size();
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}
uint64_t np2o_get_consensus(np2o_job *j, const uint32_t **pos, const uint8_t **base) {
    *pos = j->d_cns_pos.data();
    *base = j->d_cns_base.data();
    return j->d_cns_pos.size();
}
double np2o_get_seconds(np2o_job *j) { return j->seconds; }
void np2o_order_exposure(uint64_t out[4], int reset) {
    for (int i = 0; i < 4; i++) {
        out[i] = g_order_exposure[i].load();
        if (reset) g_order_exposure[i] = 0;
    }
}

// main.rs:607-645
uint64_t np2o_format_fasta(const char *tid, const uint32_t *pos, const uint8_t *base, uint64_t n, int uppercase,
                           int out_pos, uint8_t *out, uint64_t cap) {
    std::string s;
    auto up = [&](uint8_t c) -> char { return (char)(uppercase && c >= 'a' && c <= 'z' ? c - 32 : c); };
    if (out_pos) {
        for (uint64_t i = 0; i < n; i++) {
            s += tid;
            s += '\t';
            s += up(base[i]);
            s += '\t';
            s += std::to_string(pos[i]);
            s += '\n';
        }
    } else {
        if (n == 0) return 0;  // reference: first().unwrap() panics
        s += '>';
        s += tid;
        s += " start:" + std::to_string(pos[0]) + " end:" + std::to_string(pos[n - 1]) + "\n";
        for (uint64_t i = 0; i < n; i++) s += up(base[i]);
        s += '\n';
    }
    if (s.size() <= cap) memcpy(out, s.data(), s.size());
    return s.size();
}

}  // extern "C"
