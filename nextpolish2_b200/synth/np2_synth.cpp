/*
 * np2_synth.cpp — deterministic synthetic inputs for tests and bench.py
 * (SURVEY §8d "Synthetic inputs"): assembly contig, truth haplotypes, HiFi reads
 * whose alignment to the contig is known BY CONSTRUCTION (true edit script ->
 * CIGAR, indels left-aligned), emitted as raw BAM alignment records; yak table
 * synthesis (every truth-haplotype k-mer with a Poisson count) in the yak dump
 * layout (yak/htab.c:190-211); BGZF/BAM/BAI writers for the CLI tests.
 *
 * Host-only tooling.  It does not implement any part of the polish path and
 * shares no code with libnp2gpu or the oracle (the yak hash below is needed to
 * produce table keys and is restated from yak/yak-priv.h:10-38).
 */
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include <zlib.h>

namespace {

struct Rng {  // splitmix64-seeded xoshiro256**
    uint64_t s[4];
    static uint64_t sm(uint64_t &x) {
        uint64_t z = (x += 0x9E3779B97F4A7C15ULL);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    }
    explicit Rng(uint64_t seed) {
        for (auto &v : s) v = sm(seed);
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {
        uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0];
        s[3] ^= s[1];
        s[1] ^= s[2];
        s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl(s[3], 45);
        return r;
    }
    double uni() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
    uint64_t below(uint64_t n) { return n ? next() % n : 0; }
    double normal() {
        double u1 = uni(), u2 = uni();
        if (u1 < 1e-300) u1 = 1e-300;
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    }
    uint32_t poisson(double lam) {
        if (lam > 30) {
            double v = lam + std::sqrt(lam) * normal() + 0.5;
            return v < 0 ? 0 : (uint32_t)v;
        }
        double L = std::exp(-lam), p = 1;
        uint32_t k = 0;
        do {
            k++;
            p *= uni();
        } while (p > L);
        return k - 1;
    }
};

const char ACGT[] = "ACGT";

struct Var {  // a difference between a truth haplotype and the assembly, in assembly coordinates
    uint32_t pos;       // SNV/DEL: first affected assembly base; INS: inserted before assembly base `pos`
    uint8_t type;       // 0 SNV, 1 INS (haplotype has extra bases), 2 DEL (haplotype lacks assembly bases)
    uint16_t len;       // INS/DEL length
    char bases[8];      // SNV: 1 base; INS: len bases
};

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
inline int nt4(uint8_t c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return 4;
    }
}

// canonical hashed k-mers as yak count produces them (yak/count.c:28-74)
void seq_hashes(const uint8_t *seq, uint64_t len, int k, std::vector<uint64_t> &out) {
    if (k < 32) {
        uint64_t x[2] = {0, 0}, mask = (1ULL << k * 2) - 1, shift = (k - 1) * 2;
        int l = 0;
        for (uint64_t i = 0; i < len; i++) {
            int c = nt4(seq[i]);
            if (c < 4) {
                x[0] = (x[0] << 2 | c) & mask;
                x[1] = x[1] >> 2 | (uint64_t)(3 - c) << shift;
                if (++l >= k) out.push_back(yak_hash64(x[0] < x[1] ? x[0] : x[1], mask));
            } else l = 0, x[0] = x[1] = 0;
        }
    } else {
        uint64_t x[4] = {0, 0, 0, 0}, mask = (1ULL << k) - 1, shift = k - 1;
        int l = 0;
        for (uint64_t i = 0; i < len; i++) {
            int c = nt4(seq[i]);
            if (c < 4) {
                x[0] = (x[0] << 1 | (c & 1)) & mask;
                x[1] = (x[1] << 1 | (c >> 1)) & mask;
                x[2] = x[2] >> 1 | (uint64_t)(1 - (c & 1)) << shift;
                x[3] = x[3] >> 1 | (uint64_t)(1 - (c >> 1)) << shift;
                if (++l >= k) {
                    int j = x[1] < x[3] ? 0 : 1;
                    out.push_back(yak_hash64_64(x[j << 1]) + yak_hash64_64(x[j << 1 | 1]));
                }
            } else l = 0, x[0] = x[1] = x[2] = x[3] = 0;
        }
    }
}

struct Col {
    char t, q;  // '-' for gap
};

// left-align indels across exact-match columns (what minimap2/samtools normalisation gives)
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
        // gap run [i, j); shift left while the column before is an exact match equal to the run's last base
        size_t s = i, e = j;
        while (s > 1 && c[s - 1].t == c[s - 1].q && c[s - 1].t != '-') {
            char last = ins ? c[e - 1].q : c[e - 1].t;
            if (last != c[s - 1].t) break;
            // slide the gap one column to the left: only the gapped side moves, the other side keeps its bases
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

struct ReadCfg {
    double err;          // per-base read error rate
    double mean_len, sd_len, min_len, max_len;
    double frac_clip;    // fraction of reads given a long soft clip (exercise is_clip)
    double frac_lowq;    // fraction with MAPQ 0
    double frac_supp;    // fraction flagged supplementary
    int use_eqx;         // emit =/X instead of M
};

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

// one read: walk the assembly from a_s, applying the haplotype's variants then read errors
void make_read(const uint8_t *A, uint32_t L, const std::vector<Var> &vars, uint32_t a_s, uint32_t want_len,
               const ReadCfg &cfg, Rng &rng, int32_t ref_id, uint64_t read_no, std::vector<uint8_t> &out) {
    std::vector<Col> cols;
    cols.reserve(want_len + want_len / 50 + 64);
    size_t vi = std::lower_bound(vars.begin(), vars.end(), a_s,
                                 [](const Var &v, uint32_t p) { return v.pos < p; }) - vars.begin();
    uint32_t p = a_s;
    uint32_t qlen = 0;
    auto emit_hap_base = [&](char tb, char hb) {
        // hb: haplotype base aligned to assembly base tb ('-' if hb is an insertion relative to A)
        double u = rng.uni();
        if (u < cfg.err) {
            double w = rng.uni();
            if (w < 0.8) {  // homopolymer-length error: duplicate or drop this base
                if (rng.next() & 1) {
                    cols.push_back(Col{tb, hb});
                    cols.push_back(Col{'-', hb});
                    qlen += 2;
                } else {
                    if (tb != '-') cols.push_back(Col{tb, '-'});
                }
            } else if (w < 0.9) {  // substitution
                char nb = ACGT[(nt4(hb) + 1 + rng.below(3)) & 3];
                cols.push_back(Col{tb, nb});
                qlen++;
            } else {  // random 1-bp indel
                if (rng.next() & 1) {
                    cols.push_back(Col{tb, hb});
                    cols.push_back(Col{'-', ACGT[rng.below(4)]});
                    qlen += 2;
                } else {
                    if (tb != '-') cols.push_back(Col{tb, '-'});
                }
            }
        } else {
            cols.push_back(Col{tb, hb});
            qlen++;
        }
    };
    while (p < L && qlen < want_len) {
        if (vi < vars.size() && vars[vi].pos == p) {
            const Var &v = vars[vi++];
            if (v.type == 0) {
                emit_hap_base((char)A[p], v.bases[0]);
                p++;
            } else if (v.type == 1) {
                for (int x = 0; x < v.len; x++) emit_hap_base('-', v.bases[x]);
                // the assembly base at p is emitted by the next loop turn (no variant may share pos with an INS)
                emit_hap_base((char)A[p], (char)A[p]);
                p++;
            } else {
                for (int x = 0; x < v.len && p < L; x++) {
                    cols.push_back(Col{(char)A[p], '-'});
                    p++;
                }
            }
            continue;
        }
        emit_hap_base((char)A[p], (char)A[p]);
        p++;
    }
    // strip non-match columns at both ends, then left-align
    size_t b = 0, e = cols.size();
    uint32_t pos = a_s;
    while (b < e && !(cols[b].t == cols[b].q)) {
        if (cols[b].t != '-') pos++;
        b++;
    }
    while (e > b && !(cols[e - 1].t == cols[e - 1].q)) e--;
    if (e - b < 16) return;
    std::vector<Col> c(cols.begin() + b, cols.begin() + e);
    left_align(c);

    // flags / clipping / mapq
    uint16_t flag = (rng.next() & 1) ? 0x10 : 0;
    uint8_t mapq = 60;
    double u = rng.uni();
    uint32_t clip5 = 0, clip3 = 0;
    if (u < cfg.frac_clip) {
        clip5 = 150 + (uint32_t)rng.below(400);
        if (rng.next() & 1) clip3 = 150 + (uint32_t)rng.below(400);
    } else if (u < cfg.frac_clip + cfg.frac_lowq) {
        mapq = (uint8_t)rng.below(2);
    } else if (u < cfg.frac_clip + cfg.frac_lowq + cfg.frac_supp) {
        flag |= 0x800;
    }

    // CIGAR + SEQ
    std::vector<uint32_t> cigar;
    std::string seq;
    auto push_op = [&](uint32_t op, uint32_t len) {
        if (!len) return;
        if (!cigar.empty() && (cigar.back() & 15) == op) cigar.back() += len << 4;
        else cigar.push_back(len << 4 | op);
    };
    for (uint32_t x = 0; x < clip5; x++) seq.push_back(ACGT[rng.below(4)]);
    push_op(4, clip5);
    uint32_t rspan = 0;
    for (auto &cl : c) {
        if (cl.t == '-') {
            push_op(1, 1);
            seq.push_back(cl.q);
        } else if (cl.q == '-') {
            push_op(2, 1);
            rspan++;
        } else {
            push_op(cfg.use_eqx ? (cl.t == cl.q ? 7 : 8) : 0, 1);
            seq.push_back(cl.q);
            rspan++;
        }
    }
    for (uint32_t x = 0; x < clip3; x++) seq.push_back(ACGT[rng.below(4)]);
    push_op(4, clip3);

    char name[32];
    int ln = snprintf(name, sizeof name, "r%llu", (unsigned long long)read_no) + 1;
    const uint32_t l_seq = (uint32_t)seq.size();
    const uint32_t block = 32 + ln + 4 * (uint32_t)cigar.size() + (l_seq + 1) / 2 + l_seq;
    put32(out, block);
    put32(out, (uint32_t)ref_id);
    put32(out, pos);
    out.push_back((uint8_t)ln);
    out.push_back(mapq);
    put16(out, (uint16_t)reg2bin(pos, pos + rspan));
    put16(out, (uint16_t)cigar.size());
    put16(out, flag);
    put32(out, l_seq);
    put32(out, (uint32_t)-1);
    put32(out, (uint32_t)-1);
    put32(out, 0);
    out.insert(out.end(), name, name + ln);
    for (uint32_t cg : cigar) put32(out, cg);
    static const uint8_t ENC[5] = {1, 2, 4, 8, 15};
    for (uint32_t i = 0; i < l_seq; i += 2) {
        uint8_t hi = ENC[nt4(seq[i])], lo = i + 1 < l_seq ? ENC[nt4(seq[i + 1])] : 0;
        out.push_back((uint8_t)(hi << 4 | lo));
    }
    out.insert(out.end(), l_seq, 0xFF);
}

void apply_vars(const uint8_t *A, uint32_t L, const std::vector<Var> &vars, std::vector<uint8_t> &hap) {
    hap.clear();
    hap.reserve(L + L / 100);
    size_t vi = 0;
    uint32_t p = 0;
    while (p < L) {
        if (vi < vars.size() && vars[vi].pos == p) {
            const Var &v = vars[vi++];
            if (v.type == 0) {
                hap.push_back((uint8_t)v.bases[0]);
                p++;
            } else if (v.type == 1) {
                for (int x = 0; x < v.len; x++) hap.push_back((uint8_t)v.bases[x]);
                hap.push_back(A[p]);
                p++;
            } else {
                p += v.len;
            }
            continue;
        }
        hap.push_back(A[p]);
        p++;
    }
}

// draws variants at `rate` per base, min spacing 12 bp, never in the first/last 64 bp
void draw_vars(const uint8_t *A, uint32_t L, double rate, double f_hp, double f_snv, int max_indel, Rng &rng,
               std::vector<Var> &out) {
    if (rate <= 0 || L < 256) return;
    double mean_gap = 1.0 / rate;
    double p = 64 + -std::log(1 - rng.uni()) * mean_gap;
    while (p < (double)L - 64) {
        uint32_t pos = (uint32_t)p;
        Var v{};
        v.pos = pos;
        double u = rng.uni();
        if (u < f_hp) {  // homopolymer +-1 (assembly-error style)
            if (rng.next() & 1) {
                v.type = 1;
                v.len = 1;
                v.bases[0] = (char)A[pos];
            } else {
                v.type = 2;
                v.len = 1;
            }
        } else if (u < f_hp + f_snv) {
            v.type = 0;
            v.len = 1;
            v.bases[0] = ACGT[(nt4(A[pos]) + 1 + rng.below(3)) & 3];
        } else {
            int len = 1 + (int)rng.below(max_indel);
            if (rng.next() & 1) {
                v.type = 1;
                v.len = (uint16_t)len;
                for (int x = 0; x < len; x++) v.bases[x] = ACGT[rng.below(4)];
            } else {
                v.type = 2;
                v.len = (uint16_t)len;
            }
        }
        out.push_back(v);
        p += 12 + -std::log(1 - rng.uni()) * mean_gap;
    }
}

void merge_vars(const std::vector<Var> &a, const std::vector<Var> &b, std::vector<Var> &out) {
    out.clear();
    size_t i = 0, j = 0;
    int64_t last_end = -100;
    auto push = [&](const Var &v) {
        if ((int64_t)v.pos < last_end + 12) return;  // keep variants apart
        out.push_back(v);
        last_end = v.pos + (v.type == 2 ? v.len : 1);
    };
    while (i < a.size() || j < b.size()) {
        if (j >= b.size() || (i < a.size() && a[i].pos <= b[j].pos)) push(a[i++]);
        else push(b[j++]);
    }
}

struct Buf {
    std::vector<uint8_t> v;
};

}  // namespace

extern "C" {

/* i.i.d. genome with GC content `gc`, plus `tandem_frac` of the length in tandem-repeat blocks (unit 2-60 bp). */
void np2s_genome(uint64_t seed, uint32_t len, double gc, double tandem_frac, uint8_t *out) {
    Rng rng(seed);
    for (uint32_t i = 0; i < len; i++) {
        double u = rng.uni();
        bool isgc = u < gc;
        out[i] = isgc ? ((rng.next() & 1) ? 'G' : 'C') : ((rng.next() & 1) ? 'A' : 'T');
    }
    if (tandem_frac > 0) {
        uint64_t target = (uint64_t)(tandem_frac * len), done = 0;
        while (done < target) {
            uint32_t unit = 2 + (uint32_t)rng.below(59);
            uint32_t blk = 1000 + (uint32_t)rng.below(19000);
            if (blk + 2000 >= len) blk = len / 4;
            uint32_t s = 1000 + (uint32_t)rng.below(len - blk - 2000);
            for (uint32_t x = unit; x < blk; x++) out[s + x] = out[s + x % unit];
            done += blk;
        }
    }
}

/*
 * Builds one contig's worth of input.
 *   A/L           assembly contig
 *   asm_err_rate  rate of assembly errors (differences truth-hap1 vs A)
 *   het_rate      rate of extra het differences carried by hap2 (0 = haploid)
 * Returns an opaque handle; sizes via np2s_contig_sizes, data via np2s_contig_copy.
 */
struct np2s_contig {
    std::vector<uint8_t> hap1, hap2, bam;
    uint64_t n_reads = 0;
};

np2s_contig *np2s_contig_make(uint64_t seed, const uint8_t *A, uint32_t L, int32_t ref_id, double depth,
                              double asm_err_rate, double het_rate, double read_err, double mean_len, double sd_len,
                              double min_len, double max_len, double frac_clip, double frac_lowq, double frac_supp,
                              int use_eqx, int n_threads) {
    np2s_contig *c = new np2s_contig();
    Rng rng(seed);
    std::vector<Var> v1, vhet, v2;
    draw_vars(A, L, asm_err_rate, 0.7, 0.2, 3, rng, v1);
    draw_vars(A, L, het_rate, 0.0, 0.8, 5, rng, vhet);
    merge_vars(v1, vhet, v2);
    {
        std::vector<Var> tmp;
        merge_vars(v1, std::vector<Var>(), tmp);
        v1.swap(tmp);
    }
    apply_vars(A, L, v1, c->hap1);
    if (het_rate > 0) apply_vars(A, L, v2, c->hap2);

    ReadCfg cfg{read_err, mean_len, sd_len, min_len, max_len, frac_clip, frac_lowq, frac_supp, use_eqx};
    // read starts (sorted) and lengths
    uint64_t n_reads = (uint64_t)(depth * L / mean_len + 0.5);
    std::vector<std::pair<uint32_t, uint32_t>> rd(n_reads);
    for (auto &r : rd) {
        double len = mean_len + sd_len * rng.normal();
        len = std::min(std::max(len, min_len), max_len);
        if (len > L) len = L;
        uint32_t l = (uint32_t)len;
        // uniform over [-l/2, L - l/2) clipped, so contig ends keep coverage
        int64_t s = (int64_t)rng.below((uint64_t)L) - (int64_t)l / 2;
        if (s < 0) s = 0;
        if (s + l > L) s = (int64_t)L - l;
        r = {(uint32_t)s, l};
    }
    std::sort(rd.begin(), rd.end());
    if (n_threads < 1) n_threads = 1;
    std::vector<std::vector<uint8_t>> parts(n_threads);
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++) {
        th.emplace_back([&, t]() {
            uint64_t b = n_reads * t / n_threads, e = n_reads * (t + 1) / n_threads;
            for (uint64_t i = b; i < e; i++) {
                Rng rr(seed * 0x9E3779B97F4A7C15ULL + i * 2654435761ULL + 12345);
                const std::vector<Var> &vars = (het_rate > 0 && (rr.next() & 1)) ? v2 : v1;
                make_read(A, L, vars, rd[i].first, rd[i].second, cfg, rr, ref_id, i, parts[t]);
            }
        });
    }
    for (auto &x : th) x.join();
    // records from different threads are in start order already, but stripping can move pos: re-sort by pos (stable)
    std::vector<uint8_t> all;
    for (auto &p : parts) all.insert(all.end(), p.begin(), p.end());
    std::vector<std::pair<uint32_t, uint64_t>> idx;
    for (uint64_t off = 0; off < all.size();) {
        uint32_t bs, pos;
        memcpy(&bs, &all[off], 4);
        memcpy(&pos, &all[off + 8], 4);
        idx.push_back({pos, off});
        off += 4 + bs;
    }
    std::stable_sort(idx.begin(), idx.end(),
                     [](const std::pair<uint32_t, uint64_t> &a, const std::pair<uint32_t, uint64_t> &b) { return a.first < b.first; });
    c->bam.reserve(all.size());
    for (auto &x : idx) {
        uint32_t bs;
        memcpy(&bs, &all[x.second], 4);
        c->bam.insert(c->bam.end(), all.begin() + x.second, all.begin() + x.second + 4 + bs);
    }
    c->n_reads = idx.size();
    return c;
}
void np2s_contig_sizes(const np2s_contig *c, uint64_t *hap1, uint64_t *hap2, uint64_t *bam, uint64_t *n_reads) {
    *hap1 = c->hap1.size();
    *hap2 = c->hap2.size();
    *bam = c->bam.size();
    *n_reads = c->n_reads;
}
void np2s_contig_copy(const np2s_contig *c, uint8_t *hap1, uint8_t *hap2, uint8_t *bam) {
    if (hap1) memcpy(hap1, c->hap1.data(), c->hap1.size());
    if (hap2 && !c->hap2.empty()) memcpy(hap2, c->hap2.data(), c->hap2.size());
    if (bam) memcpy(bam, c->bam.data(), c->bam.size());
}
void np2s_contig_free(np2s_contig *c) { delete c; }

/*
 * Table synthesis: all canonical k-mers of the given sequences; count ~ Poisson(mean * multiplicity) clipped to
 * [1, 1023]; k-mers whose draw is < keep_min are dropped (mimics `yak count -b 37` keeping count >= 2).
 * Call once with out == NULL to get the number of distinct k-mers (upper bound), then with buffers.
 */
uint64_t np2s_table(uint64_t seed, uint32_t k, const uint8_t *const *seqs, const uint64_t *lens, uint32_t n_seqs,
                    double mean_count, uint32_t keep_min, uint64_t *out_hash, uint16_t *out_count, uint64_t cap) {
    std::vector<uint64_t> h;
    for (uint32_t i = 0; i < n_seqs; i++) seq_hashes(seqs[i], lens[i], (int)k, h);
    std::sort(h.begin(), h.end());
    Rng rng(seed ^ (0xABCDEFULL * k));
    uint64_t n = 0;
    for (size_t i = 0; i < h.size();) {
        size_t j = i;
        while (j < h.size() && h[j] == h[i]) j++;
        uint32_t c = rng.poisson(mean_count * (double)(j - i));
        if (c > 1023) c = 1023;
        if (c >= keep_min && c >= 1) {
            if (out_hash && n < cap) {
                out_hash[n] = h[i];
                out_count[n] = (uint16_t)c;
            }
            n++;
        }
        i = j;
    }
    return n;
}

/* The same table, built by `threads` workers: sequences are hashed in slices, hashes are partitioned by their top
 * bits, every partition is sorted and reduced on its own.  Counts are a pure function of (seed, k, hash, multiplicity),
 * so the result does not depend on the thread count (it differs from np2s_table's sequential draw).  Output is in
 * ascending hash order.  Returns the number of k-mers kept (call with out == NULL first). */
uint64_t np2s_table_mt(uint64_t seed, uint32_t k, const uint8_t *const *seqs, const uint64_t *lens, uint32_t n_seqs,
                       double mean_count, uint32_t keep_min, uint64_t *out_hash, uint16_t *out_count, uint64_t cap,
                       uint32_t threads) {
    const uint32_t T = std::max(1u, threads), P = 256;  // P partitions on the top 8 bits of the 2k-bit (or 64-bit) hash
    const int top_shift = (k < 32 ? 2 * (int)k : 64) - 8;
    struct Slice {
        uint32_t s;
        uint64_t b, e;
    };
    std::vector<Slice> slices;
    const uint64_t kSlice = 1u << 20;
    for (uint32_t i = 0; i < n_seqs; i++)
        for (uint64_t b = 0; b < lens[i]; b += kSlice) {
            const uint64_t e = std::min<uint64_t>(lens[i], b + kSlice + k - 1);  // k-mers starting inside [b, b + kSlice)
            slices.push_back({i, b, e});
        }
    std::vector<std::vector<std::vector<uint64_t>>> part(T, std::vector<std::vector<uint64_t>>(P));
    std::atomic<size_t> next(0);
    {
        std::vector<std::thread> th;
        for (uint32_t t = 0; t < T; t++)
            th.emplace_back([&, t] {
                std::vector<uint64_t> h;
                for (;;) {
                    const size_t i = next.fetch_add(1);
                    if (i >= slices.size()) break;
                    h.clear();
                    seq_hashes(seqs[slices[i].s] + slices[i].b, slices[i].e - slices[i].b, (int)k, h);
                    for (uint64_t v : h) part[t][(v >> top_shift) & (P - 1)].push_back(v);
                }
            });
        for (auto &x : th) x.join();
    }
    std::vector<std::vector<uint64_t>> keep_h(P);
    std::vector<std::vector<uint16_t>> keep_c(P);
    next = 0;
    {
        std::vector<std::thread> th;
        for (uint32_t t = 0; t < T; t++)
            th.emplace_back([&] {
                std::vector<uint64_t> h;
                for (;;) {
                    const size_t p = next.fetch_add(1);
                    if (p >= P) break;
                    h.clear();
                    for (uint32_t u = 0; u < T; u++) {
                        h.insert(h.end(), part[u][p].begin(), part[u][p].end());
                        std::vector<uint64_t>().swap(part[u][p]);
                    }
                    std::sort(h.begin(), h.end());
                    for (size_t i = 0; i < h.size();) {
                        size_t j = i;
                        while (j < h.size() && h[j] == h[i]) j++;
                        Rng rng(seed ^ (0xABCDEFULL * k) ^ (h[i] * 0x9E3779B97F4A7C15ULL));
                        uint32_t c = rng.poisson(mean_count * (double)(j - i));
                        if (c > 1023) c = 1023;
                        if (c >= keep_min && c >= 1) {
                            keep_h[p].push_back(h[i]);
                            keep_c[p].push_back((uint16_t)c);
                        }
                        i = j;
                    }
                }
            });
        for (auto &x : th) x.join();
    }
    uint64_t n = 0;
    for (uint32_t p = 0; p < P; p++) {
        if (out_hash)
            for (size_t i = 0; i < keep_h[p].size() && n + i < cap; i++) {
                out_hash[n + i] = keep_h[p][i];
                out_count[n + i] = keep_c[p][i];
            }
        n += keep_h[p].size();
    }
    return n;
}

/* yak dump writer (yak/htab.c:190-211), pre = 10 */
int np2s_write_yak(const char *path, uint32_t k, const uint64_t *hash, const uint16_t *count, uint64_t n) {
    const uint32_t pre = 10, nb = 1u << pre;
    std::vector<uint32_t> cnt(nb, 0);
    for (uint64_t i = 0; i < n; i++) cnt[hash[i] & (nb - 1)]++;
    std::vector<uint64_t> off(nb + 1, 0);
    for (uint32_t b = 0; b < nb; b++) off[b + 1] = off[b] + cnt[b];
    std::vector<uint64_t> keys(n);
    std::vector<uint64_t> cur(off.begin(), off.end() - 1);
    for (uint64_t i = 0; i < n; i++) keys[cur[hash[i] & (nb - 1)]++] = (hash[i] >> pre) << 10 | (count[i] & 1023);
    FILE *fp = fopen(path, "wb");
    if (!fp) return -1;
    fwrite("YAK\2", 1, 4, fp);
    uint32_t t[3] = {k, pre, 10};
    fwrite(t, 4, 3, fp);
    for (uint32_t b = 0; b < nb; b++) {
        uint32_t cap = 4;
        while (cap < cnt[b] * 4 / 3 + 1) cap <<= 1;
        uint32_t cs[2] = {cap, cnt[b]};
        fwrite(cs, 4, 2, fp);
        fwrite(&keys[off[b]], 8, cnt[b], fp);
    }
    fclose(fp);
    return 0;
}

/* 150-bp paired-end-like short reads (FASTA, both strands) drawn from the given sequences, for `yak count` */
int np2s_write_short_reads(const char *path, uint64_t seed, const uint8_t *const *seqs, const uint64_t *lens,
                           uint32_t n_seqs, double depth_each, uint32_t rlen, double sub_rate) {
    FILE *fp = fopen(path, "w");
    if (!fp) return -1;
    Rng rng(seed);
    std::string r(rlen, 'A');
    uint64_t id = 0;
    for (uint32_t s = 0; s < n_seqs; s++) {
        if (lens[s] < rlen) continue;
        uint64_t n = (uint64_t)(depth_each * lens[s] / rlen);
        for (uint64_t i = 0; i < n; i++) {
            uint64_t p = rng.below(lens[s] - rlen + 1);
            bool rev = rng.next() & 1;
            for (uint32_t x = 0; x < rlen; x++) {
                char c = rev ? "TGCA"[nt4(seqs[s][p + rlen - 1 - x]) & 3] : (char)seqs[s][p + x];
                if (rng.uni() < sub_rate) c = ACGT[(nt4(c) + 1 + rng.below(3)) & 3];
                r[x] = c;
            }
            fprintf(fp, ">s%llu\n%s\n", (unsigned long long)id++, r.c_str());
        }
    }
    fclose(fp);
    return 0;
}

/* ---- BGZF / BAM / BAI writers ---- */

static void bgzf_block(std::vector<uint8_t> &dst, const uint8_t *data, size_t n, int level) {
    uint8_t out[70000];
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
    zs.next_in = (Bytef *)data;
    zs.avail_in = (uInt)n;
    zs.next_out = out + 18;
    zs.avail_out = sizeof out - 18 - 8;
    deflate(&zs, Z_FINISH);
    size_t clen = zs.total_out;
    deflateEnd(&zs);
    static const uint8_t hdr[16] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 'B', 'C', 2, 0};
    memcpy(out, hdr, 16);
    uint16_t bsize = (uint16_t)(clen + 25);
    out[16] = (uint8_t)bsize;
    out[17] = (uint8_t)(bsize >> 8);
    uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), data, (uInt)n);
    uint32_t isz = (uint32_t)n;
    memcpy(out + 18 + clen, &crc, 4);
    memcpy(out + 18 + clen + 4, &isz, 4);
    dst.assign(out, out + 18 + clen + 8);
}

/*
 * Writes <path> (BAM) and <path>.bai.  names: n_ref NUL-terminated names concatenated; recs[i]/rec_len[i]: the raw
 * record blob of contig i (already coordinate-sorted, refID == i).
 */
int np2s_write_bam(const char *path, uint32_t n_ref, const char *names, const uint32_t *ref_len,
                   const uint8_t *const *recs, const uint64_t *rec_len, int level) {
    FILE *fp = fopen(path, "wb");
    if (!fp) return -1;
    std::vector<uint8_t> hdr;
    hdr.insert(hdr.end(), {'B', 'A', 'M', 1});
    std::string text = "@HD\tVN:1.6\tSO:coordinate\n";
    std::vector<std::string> nm;
    const char *q = names;
    for (uint32_t i = 0; i < n_ref; i++) {
        nm.emplace_back(q);
        q += nm.back().size() + 1;
        text += "@SQ\tSN:" + nm.back() + "\tLN:" + std::to_string(ref_len[i]) + "\n";
    }
    put32(hdr, (uint32_t)text.size());
    hdr.insert(hdr.end(), text.begin(), text.end());
    put32(hdr, n_ref);
    for (uint32_t i = 0; i < n_ref; i++) {
        put32(hdr, (uint32_t)nm[i].size() + 1);
        hdr.insert(hdr.end(), nm[i].begin(), nm[i].end());
        hdr.push_back(0);
        put32(hdr, ref_len[i]);
    }
    // Blocks are cut first (their boundaries only depend on the record sizes), compressed in parallel afterwards; until
    // then a virtual offset is (index of the block << 16 | offset inside it) and `coff` counts blocks.
    uint64_t coff = 0;  // index of the block being filled
    std::vector<uint8_t> blk;
    std::vector<std::vector<uint8_t>> raw_blocks;
    auto flush = [&]() {
        if (blk.empty()) return;
        raw_blocks.emplace_back(blk);
        coff++;
        blk.clear();
    };
    for (size_t o = 0; o < hdr.size();) {
        size_t n = std::min<size_t>(0xff00, hdr.size() - o);
        blk.assign(hdr.begin() + o, hdr.begin() + o + n);
        flush();
        o += n;
    }
    // index: per ref, bins -> chunks, 16 kb linear index
    struct RefIdx {
        std::vector<std::pair<uint32_t, std::pair<uint64_t, uint64_t>>> chunks;  // (bin, (beg, end))
        std::vector<uint64_t> lin;
    };
    std::vector<RefIdx> ridx(n_ref);
    for (uint32_t r = 0; r < n_ref; r++) {
        const uint8_t *b = recs[r];
        for (uint64_t off = 0; off < rec_len[r];) {
            uint32_t bs;
            memcpy(&bs, b + off, 4);
            if (blk.size() + 4 + bs > 0xff00) flush();
            uint64_t vbeg = coff << 16 | blk.size();
            // records larger than a block are split across blocks
            uint64_t left = 4 + bs, so = off;
            while (left) {
                size_t n = std::min<uint64_t>(left, 0xff00 - blk.size());
                blk.insert(blk.end(), b + so, b + so + n);
                so += n;
                left -= n;
                if (blk.size() >= 0xff00) flush();
            }
            uint64_t vend = coff << 16 | blk.size();
            int32_t pos;
            uint16_t bin, ncig;
            memcpy(&pos, b + off + 8, 4);
            memcpy(&bin, b + off + 14, 2);
            memcpy(&ncig, b + off + 16, 2);
            uint8_t lname = b[off + 12];
            uint32_t span = 0;
            for (uint32_t ci = 0; ci < ncig; ci++) {
                uint32_t cg;
                memcpy(&cg, b + off + 36 + lname + 4 * ci, 4);
                uint32_t op = cg & 15;
                if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) span += cg >> 4;
            }
            if (!span) span = 1;
            ridx[r].chunks.push_back({bin, {vbeg, vend}});
            for (uint32_t w = (uint32_t)pos >> 14; w <= ((uint32_t)pos + span - 1) >> 14; w++) {
                if (ridx[r].lin.size() <= w) ridx[r].lin.resize(w + 1, 0);
                if (ridx[r].lin[w] == 0) ridx[r].lin[w] = vbeg;
            }
            off += 4 + bs;
        }
    }
    flush();
    std::vector<uint64_t> block_coff(raw_blocks.size() + 1, 0);
    {
        std::vector<std::vector<uint8_t>> packed(raw_blocks.size());
        std::atomic<size_t> next{0};
        auto work = [&]() {
            for (size_t i; (i = next.fetch_add(1)) < raw_blocks.size();) {
                bgzf_block(packed[i], raw_blocks[i].data(), raw_blocks[i].size(), level);
                std::vector<uint8_t>().swap(raw_blocks[i]);
            }
        };
        const unsigned T = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
        std::vector<std::thread> th;
        for (unsigned t = 1; t < T; t++) th.emplace_back(work);
        work();
        for (auto &t : th) t.join();
        for (size_t i = 0; i < packed.size(); i++) {
            fwrite(packed[i].data(), 1, packed[i].size(), fp);
            block_coff[i + 1] = block_coff[i] + packed[i].size();
        }
    }
    auto real_voff = [&](uint64_t v) { return block_coff[v >> 16] << 16 | (v & 0xFFFF); };
    for (RefIdx &ri : ridx) {
        for (auto &c : ri.chunks) {
            c.second.first = real_voff(c.second.first);
            c.second.second = real_voff(c.second.second);
        }
        for (uint64_t &l : ri.lin)
            if (l) l = real_voff(l);
    }
    static const uint8_t eof[28] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0, 27, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    fwrite(eof, 1, 28, fp);
    fclose(fp);

    std::string bai = std::string(path) + ".bai";
    fp = fopen(bai.c_str(), "wb");
    if (!fp) return -1;
    fwrite("BAI\1", 1, 4, fp);
    fwrite(&n_ref, 4, 1, fp);
    for (uint32_t r = 0; r < n_ref; r++) {
        auto &ch = ridx[r].chunks;
        std::stable_sort(ch.begin(), ch.end(), [](const std::pair<uint32_t, std::pair<uint64_t, uint64_t>> &a,
                                                   const std::pair<uint32_t, std::pair<uint64_t, uint64_t>> &b) { return a.first < b.first; });
        uint32_t n_bin = 0;
        for (size_t i = 0; i < ch.size(); i++)
            if (i == 0 || ch[i].first != ch[i - 1].first) n_bin++;
        fwrite(&n_bin, 4, 1, fp);
        for (size_t i = 0; i < ch.size();) {
            size_t j = i;
            while (j < ch.size() && ch[j].first == ch[i].first) j++;
            uint32_t bin = ch[i].first;
            // merge adjacent chunks
            std::vector<std::pair<uint64_t, uint64_t>> m;
            for (size_t x = i; x < j; x++) {
                if (!m.empty() && ch[x].second.first <= m.back().second) m.back().second = std::max(m.back().second, ch[x].second.second);
                else m.push_back(ch[x].second);
            }
            int32_t nch = (int32_t)m.size();
            fwrite(&bin, 4, 1, fp);
            fwrite(&nch, 4, 1, fp);
            for (auto &c : m) {
                fwrite(&c.first, 8, 1, fp);
                fwrite(&c.second, 8, 1, fp);
            }
            i = j;
        }
        int32_t n_intv = (int32_t)ridx[r].lin.size();
        // fill empty windows with the next offset, as samtools does backwards-fill with previous
        for (size_t w = 1; w < ridx[r].lin.size(); w++)
            if (ridx[r].lin[w] == 0) ridx[r].lin[w] = ridx[r].lin[w - 1];
        fwrite(&n_intv, 4, 1, fp);
        fwrite(ridx[r].lin.data(), 8, ridx[r].lin.size(), fp);
    }
    fclose(fp);
    return 0;
}

}  // extern "C"
