// np2_secondary.cpp — `-S / --use_secondary` (src/utils/secondary.rs:8-158, main.rs:1775-1788), host side.
//
// Secondary alignments carry no SEQ in a BAM (`*`).  The reference makes two passes over the whole file before
// polishing: (1) the query names of every secondary record, (2) the SEQ of the PRIMARY record (neither secondary nor
// supplementary) of each such name, turned back into the read's original orientation.  While polishing, a secondary
// record that passes the filter is expanded against that sequence (reverse-complemented again when the secondary
// record is on the reverse strand).
//
// Here the same three steps are C-ABI calls over record blobs (np2_secmap_*); the last one REWRITES a contig's
// record blob so that every secondary record carries its SEQ, after which the polish path (host parse, K0..K6) needs
// nothing special: with use_secondary set it simply no longer rejects flag 0x100.
#include <cstring>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/np2gpu.h"
#include "np2_error.h"
#include "np2_host.h"

namespace np2 {

struct SecMap {
    std::unordered_set<std::string> ids;
    std::unordered_map<std::string, std::vector<uint8_t>> seqs;  // one BAM 4-bit code per byte, original orientation
};

namespace {
struct Rec {
    const uint8_t *r;  // payload (after block_size)
    uint32_t bs, l_name, n_cig, flag;
    int32_t l_seq;
};
// walks the block_size chain; f(rec) per record
template <class F>
void for_each_record(const uint8_t *bam, uint64_t len, F f) {
    uint64_t p = 0;
    while (p < len) {
        if (p + 4 > len) throw Error(NP2_ERR_FORMAT, "BAM/SAM parsing failed!");
        int32_t bs;
        memcpy(&bs, bam + p, 4);
        if (bs < 32 || p + 4 + (uint64_t)bs > len) throw Error(NP2_ERR_FORMAT, "BAM/SAM parsing failed!");
        Rec x;
        x.r = bam + p + 4;
        x.bs = (uint32_t)bs;
        x.l_name = x.r[8];
        uint16_t nc, fl;
        memcpy(&nc, x.r + 12, 2);
        memcpy(&fl, x.r + 14, 2);
        memcpy(&x.l_seq, x.r + 16, 4);
        x.n_cig = nc;
        x.flag = fl;
        if (x.l_seq < 0 || x.l_name == 0 ||
            32ull + x.l_name + 4ull * x.n_cig + ((uint64_t)x.l_seq + 1) / 2 + (uint64_t)x.l_seq > (uint64_t)bs)
            throw Error(NP2_ERR_FORMAT, "BAM/SAM parsing failed!");
        f(x);
        p += 4 + (uint64_t)bs;
    }
}
inline std::string qname(const Rec &x) {  // Record::qname(): without the trailing NUL
    return std::string((const char *)x.r + 32, x.l_name - 1);
}
// reverse_complement_seq_u8 (secondary.rs:68-83) on BAM codes: A<->T, C<->G, everything else unchanged
inline uint8_t comp4(uint8_t c) { return c == 1 ? 8 : c == 8 ? 1 : c == 2 ? 4 : c == 4 ? 2 : c; }
}  // namespace

void secmap_scan_ids(SecMap &m, const uint8_t *bam, uint64_t len) {  // retrieve_secondary_ids secondary.rs:8-66
    for_each_record(bam, len, [&](const Rec &x) {
        if (x.flag & 0x100) m.ids.insert(qname(x));
    });
}
void secmap_scan_seqs(SecMap &m, const uint8_t *bam, uint64_t len) {  // retrieve_secondary_seq_from_bam secondary.rs:85-150
    for_each_record(bam, len, [&](const Rec &x) {
        if (x.flag & 0x900) return;
        std::string q = qname(x);
        if (!m.ids.count(q)) return;
        const uint8_t *sq = x.r + 32 + x.l_name + 4ull * x.n_cig;
        std::vector<uint8_t> s((size_t)x.l_seq);
        for (int32_t i = 0; i < x.l_seq; i++) s[i] = (i & 1) ? (sq[i >> 1] & 15) : (sq[i >> 1] >> 4);
        if (x.flag & 0x10) {
            std::vector<uint8_t> t(s.size());
            for (size_t i = 0; i < s.size(); i++) t[i] = comp4(s[s.size() - 1 - i]);
            s.swap(t);
        }
        if (!m.seqs.emplace(std::move(q), std::move(s)).second)  // assert!(seqs.insert(..).is_none()) secondary.rs:131
            throw Error(NP2_ERR_FORMAT, "two primary records share the query name of a secondary alignment");
    });
}
uint64_t secmap_fill(const SecMap &m, const uint8_t *bam, uint64_t len, uint8_t *out, uint64_t cap) {
    uint64_t w = 0;
    auto put = [&](const void *p, uint64_t n) {
        if (out && w + n <= cap) memcpy(out + w, p, n);
        w += n;
    };
    for_each_record(bam, len, [&](const Rec &x) {
        if (!(x.flag & 0x100)) {
            put(x.r - 4, 4 + (uint64_t)x.bs);
            return;
        }
        auto it = m.seqs.find(qname(x));
        // no primary: the reference panics (HashMap index, main.rs:1778) if the record passes the filter; it is left
        // without SEQ so that the parse reports exactly that case
        static const std::vector<uint8_t> none;
        const std::vector<uint8_t> &s = it == m.seqs.end() ? none : it->second;
        const uint64_t head = 32ull + x.l_name + 4ull * x.n_cig;
        const uint64_t old_tail = head + ((uint64_t)x.l_seq + 1) / 2 + (uint64_t)x.l_seq;
        const uint64_t n = s.size();
        const uint32_t bs = (uint32_t)(head + (n + 1) / 2 + n + (x.bs - old_tail));
        std::vector<uint8_t> rec(4 + (size_t)bs, 0);
        memcpy(rec.data(), &bs, 4);
        memcpy(rec.data() + 4, x.r, head);
        const int32_t l_seq = (int32_t)n;
        memcpy(rec.data() + 4 + 16, &l_seq, 4);
        uint8_t *sq = rec.data() + 4 + head;
        const bool rev = x.flag & 0x10;
        for (uint64_t i = 0; i < n; i++) {
            const uint8_t c = rev ? comp4(s[n - 1 - i]) : s[i];
            sq[i >> 1] |= (i & 1) ? c : (uint8_t)(c << 4);
        }
        memset(sq + (n + 1) / 2, 0xFF, n);
        memcpy(sq + (n + 1) / 2 + n, x.r + old_tail, x.bs - old_tail);  // aux tags
        put(rec.data(), rec.size());
    });
    return w;
}

SecMap *secmap_new() { return new SecMap(); }
void secmap_delete(SecMap *m) { delete m; }
uint64_t secmap_counts(const SecMap &m, uint64_t *n_seqs) {
    if (n_seqs) *n_seqs = m.seqs.size();
    return m.ids.size();
}

}  // namespace np2
