/*
 * np2_align.cpp — a small HiFi-read-to-contig aligner for test inputs.
 *
 * BASELINE.json configs[0] is the reference's bundled test set (test/asm.fa.gz +
 * test/hifi.fasta.gz): raw reads, no BAM.  The reference README aligns them with
 * minimap2/winnowmap, which is not in this image, so the tests build their BAM
 * records with this tool: exact-k-mer anchors against the contig (unique 19-mers),
 * longest-increasing chain, unit-cost global alignment between anchors, ungapped
 * extension at the ends (the rest soft-clipped), indels left-aligned, CIGAR with M/I/D/S.
 *
 * Host-only tooling; it implements no part of the polish path (the polisher only
 * consumes the records) and is linked into libnp2synth.so.
 */
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

inline int nt4(uint8_t c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return 4;
    }
}
inline int reg2bin(int64_t beg, int64_t end) {
    --end;
    if (beg >> 14 == end >> 14) return ((1 << 15) - 1) / 7 + (beg >> 14);
    if (beg >> 17 == end >> 17) return ((1 << 12) - 1) / 7 + (beg >> 17);
    if (beg >> 20 == end >> 20) return ((1 << 9) - 1) / 7 + (beg >> 20);
    if (beg >> 23 == end >> 23) return ((1 << 6) - 1) / 7 + (beg >> 23);
    if (beg >> 26 == end >> 26) return ((1 << 3) - 1) / 7 + (beg >> 26);
    return 0;
}
void put32(std::vector<uint8_t> &b, uint32_t v) {
    for (int i = 0; i < 4; i++) b.push_back((uint8_t)(v >> (8 * i)));
}
void put16(std::vector<uint8_t> &b, uint16_t v) {
    b.push_back((uint8_t)v);
    b.push_back((uint8_t)(v >> 8));
}

constexpr int K = 19;

struct Index {
    std::unordered_map<uint64_t, int32_t> pos;  // k-mer -> contig position, -1 if it occurs more than once
    void build(const uint8_t *t, uint32_t L) {
        pos.reserve(L * 2);
        uint64_t x = 0, mask = (1ULL << 2 * K) - 1;
        int l = 0;
        for (uint32_t i = 0; i < L; i++) {
            int c = nt4(t[i]);
            if (c > 3) {
                l = 0;
                x = 0;
                continue;
            }
            x = (x << 2 | (uint64_t)c) & mask;
            if (++l >= K) {
                auto it = pos.find(x);
                if (it == pos.end()) pos.emplace(x, (int32_t)(i + 1 - K));
                else it->second = -1;
            }
        }
    }
};

struct Anchor {
    int32_t q, t;
};

void find_anchors(const Index &ix, const std::vector<uint8_t> &q, std::vector<Anchor> &out) {
    out.clear();
    uint64_t x = 0, mask = (1ULL << 2 * K) - 1;
    int l = 0;
    for (size_t i = 0; i < q.size(); i++) {
        int c = nt4(q[i]);
        if (c > 3) {
            l = 0;
            x = 0;
            continue;
        }
        x = (x << 2 | (uint64_t)c) & mask;
        if (++l >= K) {
            auto it = ix.pos.find(x);
            if (it != ix.pos.end() && it->second >= 0) out.push_back(Anchor{(int32_t)(i + 1 - K), it->second});
        }
    }
}

// longest chain with strictly increasing contig position (anchors arrive sorted by read position)
void chain(std::vector<Anchor> &a) {
    const size_t n = a.size();
    if (!n) return;
    std::vector<int32_t> tail, tail_idx, prev(n, -1);
    for (size_t i = 0; i < n; i++) {
        size_t p = std::lower_bound(tail.begin(), tail.end(), a[i].t) - tail.begin();
        if (p == tail.size()) {
            tail.push_back(a[i].t);
            tail_idx.push_back((int32_t)i);
        } else {
            tail[p] = a[i].t;
            tail_idx[p] = (int32_t)i;
        }
        prev[i] = p ? tail_idx[p - 1] : -1;
    }
    std::vector<Anchor> c;
    for (int32_t i = tail_idx.back(); i >= 0; i = prev[i]) c.push_back(a[i]);
    std::reverse(c.begin(), c.end());
    a.swap(c);
}

struct Col {
    char t, q;
};

// unit-cost global alignment of q[0,n) against t[0,m); appends columns
bool global_align(const uint8_t *q, int n, const uint8_t *t, int m, std::vector<Col> &cols) {
    if (n == 0) {
        for (int j = 0; j < m; j++) cols.push_back(Col{(char)t[j], '-'});
        return true;
    }
    if (m == 0) {
        for (int i = 0; i < n; i++) cols.push_back(Col{'-', (char)q[i]});
        return true;
    }
    if ((int64_t)n * m > 16000000) return false;
    std::vector<int32_t> D((size_t)(n + 1) * (m + 1));
    auto at = [&](int i, int j) -> int32_t & { return D[(size_t)i * (m + 1) + j]; };
    for (int j = 0; j <= m; j++) at(0, j) = j;
    for (int i = 1; i <= n; i++) {
        at(i, 0) = i;
        for (int j = 1; j <= m; j++) {
            int32_t s = at(i - 1, j - 1) + (q[i - 1] == t[j - 1] ? 0 : 1);
            int32_t d = at(i, j - 1) + 1, x = at(i - 1, j) + 1;
            at(i, j) = std::min(s, std::min(d, x));
        }
    }
    std::vector<Col> rev;
    int i = n, j = m;
    while (i > 0 || j > 0) {
        if (i > 0 && j > 0 && at(i, j) == at(i - 1, j - 1) + (q[i - 1] == t[j - 1] ? 0 : 1)) {
            rev.push_back(Col{(char)t[j - 1], (char)q[i - 1]});
            i--, j--;
        } else if (j > 0 && at(i, j) == at(i, j - 1) + 1) {
            rev.push_back(Col{(char)t[j - 1], '-'});
            j--;
        } else {
            rev.push_back(Col{'-', (char)q[i - 1]});
            i--;
        }
    }
    cols.insert(cols.end(), rev.rbegin(), rev.rend());
    return true;
}

void left_align(std::vector<Col> &c) {
    const size_t n = c.size();
    size_t i = 0;
    while (i < n) {
        if (c[i].t != '-' && c[i].q != '-') {
            i++;
            continue;
        }
        const bool ins = c[i].t == '-';
        size_t j = i;
        while (j < n && (ins ? (c[j].t == '-') : (c[j].q == '-'))) j++;
        size_t s = i, e = j;
        while (s > 1 && c[s - 1].t == c[s - 1].q) {
            char last = ins ? c[e - 1].q : c[e - 1].t;
            if (last != c[s - 1].t) break;
            Col m = c[s - 1];
            if (ins) {
                c[s - 1].t = '-';
                c[e - 1].t = m.t;
            } else {
                c[s - 1].q = '-';
                c[e - 1].q = m.q;
            }
            s--;
            e--;
        }
        i = j;
    }
}

void revcomp(std::vector<uint8_t> &s) {
    std::reverse(s.begin(), s.end());
    for (auto &c : s) {
        switch (c) {
            case 'A': case 'a': c = 'T'; break;
            case 'C': case 'c': c = 'G'; break;
            case 'G': case 'g': c = 'C'; break;
            case 'T': case 't': c = 'A'; break;
            default: break;
        }
    }
}

struct Rec {
    uint32_t pos;
    std::vector<uint8_t> bytes;
};

bool align_read(const Index &ix, const uint8_t *T, uint32_t L, const uint8_t *seq, uint32_t len, const char *name,
                int32_t ref_id, Rec &out) {
    std::vector<uint8_t> q(seq, seq + len), qr(seq, seq + len);
    for (auto &c : q) c = (uint8_t)toupper(c);
    qr = q;
    revcomp(qr);
    std::vector<Anchor> af, ar;
    find_anchors(ix, q, af);
    find_anchors(ix, qr, ar);
    chain(af);
    chain(ar);
    const bool rev = ar.size() > af.size();
    std::vector<Anchor> &a = rev ? ar : af;
    const std::vector<uint8_t> &r = rev ? qr : q;
    if (a.size() < 20) return false;

    std::vector<Col> cols;
    cols.reserve(len + len / 50);
    // ungapped extension to the left of the first anchor
    int32_t qi = a[0].q, tj = a[0].t;
    while (qi > 0 && tj > 0 && r[qi - 1] == T[tj - 1]) qi--, tj--;
    const int32_t q_start = qi, t_start = tj;
    for (const Anchor &an : a) {
        if (an.q < qi || an.t < tj) {
            // overlaps what is already aligned: usable only when it continues the current diagonal
            if (an.q - an.t == qi - tj && an.q + K > qi) {
                for (int32_t x = qi; x < an.q + K; x++) cols.push_back(Col{(char)T[tj + (x - qi)], (char)r[x]});
                tj += an.q + K - qi;
                qi = an.q + K;
            }
            continue;
        }
        if (!global_align(&r[qi], an.q - qi, &T[tj], an.t - tj, cols)) return false;
        for (int x = 0; x < K; x++) cols.push_back(Col{(char)T[an.t + x], (char)r[an.q + x]});
        qi = an.q + K;
        tj = an.t + K;
    }
    while (qi < (int32_t)len && tj < (int32_t)L && r[qi] == T[tj]) {
        cols.push_back(Col{(char)T[tj], (char)r[qi]});
        qi++, tj++;
    }
    left_align(cols);

    std::vector<uint32_t> cigar;
    auto push_op = [&](uint32_t op, uint32_t n) {
        if (!n) return;
        if (!cigar.empty() && (cigar.back() & 15) == op) cigar.back() += n << 4;
        else cigar.push_back(n << 4 | op);
    };
    push_op(4, (uint32_t)q_start);
    uint32_t rspan = 0;
    for (auto &cl : cols) {
        if (cl.t == '-') push_op(1, 1);
        else if (cl.q == '-') push_op(2, 1), rspan++;
        else push_op(0, 1), rspan++;
    }
    push_op(4, len - (uint32_t)qi);
    if (cigar.size() > 65535) return false;

    std::vector<uint8_t> &b = out.bytes;
    const int ln = (int)strlen(name) + 1;
    const uint32_t block = 32 + ln + 4 * (uint32_t)cigar.size() + (len + 1) / 2 + len;
    out.pos = (uint32_t)t_start;
    put32(b, block);
    put32(b, (uint32_t)ref_id);
    put32(b, (uint32_t)t_start);
    b.push_back((uint8_t)ln);
    b.push_back(60);
    put16(b, (uint16_t)reg2bin(t_start, t_start + rspan));
    put16(b, (uint16_t)cigar.size());
    put16(b, rev ? 0x10 : 0);
    put32(b, len);
    put32(b, (uint32_t)-1);
    put32(b, (uint32_t)-1);
    put32(b, 0);
    b.insert(b.end(), name, name + ln);
    for (uint32_t cg : cigar) put32(b, cg);
    static const uint8_t ENC[5] = {1, 2, 4, 8, 15};
    for (uint32_t i = 0; i < len; i += 2) {
        uint8_t hi = ENC[nt4(r[i])], lo = i + 1 < len ? ENC[nt4(r[i + 1])] : 0;
        b.push_back((uint8_t)(hi << 4 | lo));
    }
    b.insert(b.end(), len, 0xFF);
    return true;
}

}  // namespace

extern "C" {

struct np2s_aln {
    std::vector<uint8_t> bam;
    uint64_t n_aligned = 0;
};

/* Aligns n reads (concatenated bytes + n+1 offsets; names NUL-separated) to the contig and returns the BAM alignment
 * records sorted by position (stable: ties keep read order). */
np2s_aln *np2s_align(const uint8_t *T, uint32_t L, int32_t ref_id, const uint8_t *seqs, const uint64_t *off,
                     const char *names, uint64_t n, int n_threads) {
    Index ix;
    ix.build(T, L);
    std::vector<const char *> nm(n);
    const char *p = names;
    for (uint64_t i = 0; i < n; i++) {
        nm[i] = p;
        p += strlen(p) + 1;
    }
    std::vector<Rec> recs(n);
    std::vector<uint8_t> ok(n, 0);
    if (n_threads < 1) n_threads = 1;
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++)
        th.emplace_back([&, t]() {
            for (uint64_t i = t; i < n; i += n_threads)
                ok[i] = align_read(ix, T, L, seqs + off[i], (uint32_t)(off[i + 1] - off[i]), nm[i], ref_id, recs[i]);
        });
    for (auto &x : th) x.join();
    std::vector<uint64_t> order;
    for (uint64_t i = 0; i < n; i++)
        if (ok[i]) order.push_back(i);
    std::stable_sort(order.begin(), order.end(), [&](uint64_t a, uint64_t b) { return recs[a].pos < recs[b].pos; });
    np2s_aln *r = new np2s_aln();
    for (uint64_t i : order) r->bam.insert(r->bam.end(), recs[i].bytes.begin(), recs[i].bytes.end());
    r->n_aligned = order.size();
    return r;
}
uint64_t np2s_aln_size(const np2s_aln *a, uint64_t *n_aligned) {
    if (n_aligned) *n_aligned = a->n_aligned;
    return a->bam.size();
}
void np2s_aln_copy(const np2s_aln *a, uint8_t *out) { memcpy(out, a->bam.data(), a->bam.size()); }
void np2s_aln_free(np2s_aln *a) { delete a; }
}
