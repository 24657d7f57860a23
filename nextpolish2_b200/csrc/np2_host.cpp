// np2_host.cpp — host phases of the polish path (see np2_host.h).  Citations are to the reference
// (Nextomics/NextPolish2 @ 283dc5a).  Nothing here calls into oracle/.
#include "np2_host.h"

#include "np2_error.h"

#include <algorithm>
#include <cstring>
#include <map>
#include <set>
#include <stdexcept>
#include <unordered_map>

namespace np2 {

[[noreturn]] static void herr(int code, const std::string &m) { throw Error(code, m); }

/* ================================================================= ingest */

void parse_records(const uint8_t *bam, uint64_t bam_len, uint32_t tlen, const np2_opts &opt, Ingest &out) {
    out = Ingest();
    out.op_off.push_back(0);
    out.nib_off.push_back(0);
    out.ck_off.push_back(0);
    uint64_t off = 0;
    int32_t rec = -1;
    while (off + 4 <= bam_len) {
        rec++;
        int32_t bs;
        memcpy(&bs, bam + off, 4);
        if (bs < 32 || off + 4 + (uint64_t)bs > bam_len) herr(NP2_ERR_FORMAT, "BAM/SAM parsing failed!");
        const uint8_t *r = bam + off + 4;
        const uint64_t rec_off = off + 4;
        off += 4 + (uint64_t)bs;
        int32_t ref_id, pos, l_seq;
        uint16_t n_cig, flag;
        memcpy(&ref_id, r, 4);
        memcpy(&pos, r + 4, 4);
        const uint32_t l_name = r[8], mapq = r[9];
        memcpy(&n_cig, r + 12, 2);
        memcpy(&flag, r + 14, 2);
        memcpy(&l_seq, r + 16, 4);
        if (l_seq < 0 || 32ull + l_name + 4ull * n_cig + ((uint64_t)l_seq + 1) / 2 > (uint64_t)bs)
            herr(NP2_ERR_FORMAT, "BAM/SAM parsing failed!");
        out.all_tid.push_back(ref_id);
        out.all_pos.push_back(pos);
        const uint8_t *cg = r + 32 + l_name;
        // seq_len_from_cigar(true), bam_endpos
        uint64_t rlen = 0, rspan = 0;
        for (uint32_t i = 0; i < n_cig; i++) {
            uint32_t c;
            memcpy(&c, cg + 4 * i, 4);
            const uint32_t l = c >> 4, op = c & 15;
            if (op == 0 || op == 1 || op == 4 || op == 5 || op == 7 || op == 8) rlen += l;
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rspan += l;
        }
        const int64_t span = ((flag & 4) || n_cig == 0 || rspan == 0) ? 1 : (int64_t)rspan;
        const int64_t need = std::max<int64_t>((int64_t)opt.min_map_len, (int64_t)((float)rlen * opt.min_map_fra));
        if ((flag & 0x404) || (int16_t)mapq <= (int16_t)opt.min_map_qual || rlen <= opt.min_read_len ||
            ((flag & 0x100) && !opt.use_secondary) || ((flag & 0x800) && !opt.use_supplementary) || span < need)
            continue;
        if (pos < 0 || (uint64_t)pos > tlen) herr(NP2_ERR_FORMAT, "alignment starts outside the contig");
        // fill_with_cigar bookkeeping (main.rs:386-440) without materialising the strings
        uint32_t qs = 0, ts = 0, col = 0, aln_q_s = 0, aln_q_e = 0;
        bool first = true;
        for (uint32_t i = 0; i < n_cig; i++) {
            uint32_t c;
            memcpy(&c, cg + 4 * i, 4);
            const uint32_t l = c >> 4, op = c & 15;
            switch (op) {
                case 4:
                    qs += l;
                    if (first) aln_q_s = qs;
                    else aln_q_e = qs - l;
                    break;
                case 0: case 7: case 8: case 1: case 2:
                    if (op != 2 && (uint64_t)qs + l > (uint64_t)l_seq)
                        herr(NP2_ERR_FORMAT, "CIGAR consumes more query bases than SEQ holds");
                    if (op != 1 && (uint64_t)pos + ts + l > tlen)
                        herr(NP2_ERR_FORMAT, "alignment runs past the end of the contig");
                    if (l) {
                        out.op_col.push_back(col);
                        out.op_q.push_back(qs);
                        out.op_t.push_back(ts);
                        out.op_cig.push_back(c);
                    }
                    col += l;
                    if (op != 2) qs += l;
                    if (op != 1) ts += l;
                    break;
                case 5:
                    break;
                default:
                    herr(NP2_ERR_FORMAT, "Unknown cigar");
            }
            first = false;
        }
        if (aln_q_e == 0) aln_q_e = qs;
        out.rec_idx.push_back(rec);
        out.pos.push_back((uint32_t)pos);
        out.ncols.push_back(col);
        out.rlen.push_back((uint32_t)rlen);
        out.is_clip.push_back((uint32_t)(aln_q_e - aln_q_s + opt.max_clip_len) < (uint32_t)rlen ? 1 : 0);  // main.rs:1796
        out.seq_off.push_back(rec_off + 32 + l_name + 4ull * n_cig);
        out.op_off.push_back((uint32_t)out.op_col.size());
        out.nib_off.push_back(out.nib_off.back() + ((((uint64_t)col / 16 + 1) * 8 + 15) & ~15ull));
        out.ck_off.push_back(out.ck_off.back() + (col + 31) / 32);
        out.total_cols += col;
    }
    if (off != bam_len) herr(NP2_ERR_FORMAT, "BAM/SAM parsing failed!");
}

/* ================================================================= regions */

void find_regions(const uint32_t *cpos, const uint8_t *cbase, const uint8_t *cflags, uint64_t n, const uint32_t *events,
                  uint64_t n_events, Regions &out) {
    out.start.clear();
    out.end.clear();
    if (n == 0) return;
    // reversed indexing: rp = n - 1 - i is the index the reference uses while backtracking (main.rs:1570-1626)
    auto P = [&](uint64_t rp) { return cpos[n - 1 - rp]; };
    auto B = [&](uint64_t rp) { return cbase[n - 1 - rp]; };
    const uint64_t NONE = UINT64_MAX;
    bool has_lq = false;
    uint64_t lq_s = NONE, lq_e = 0;
    auto try_close = [&](uint64_t from, uint64_t to) {  // HQ bases rp in [from, to)
        for (uint64_t rp = std::max(from, lq_e + 5); rp < to; rp++) {
            if (P(rp - 1) != P(rp - 2) && B(rp - 1) != B(rp - 2)) {
                lq_e = rp - 2;
                lq_s = lq_s > 2 ? lq_s - 2 : 1;
                while (lq_s > 1 && (P(lq_s - 1) == P(lq_s) || B(lq_s - 1) == B(lq_s))) lq_s--;
                if (!out.start.empty() && P(lq_s) >= out.start.back()) {
                    out.start.back() = P(lq_e);
                } else {
                    out.end.push_back(P(lq_s));
                    out.start.push_back(P(lq_e));
                }
                has_lq = false;
                lq_s = NONE;
                return;
            }
        }
    };
    uint64_t next_rp = 0;
    for (uint64_t e = n_events; e-- > 0;) {
        const uint64_t rp = n - 1 - events[e];
        if (has_lq) try_close(next_rp, rp);
        const uint8_t f = cflags[events[e]];
        if (f & 2) {
            has_lq = false;
            lq_s = NONE;
        } else if (f & 1) {
            if (lq_s == NONE) lq_s = rp;
            lq_e = rp;
            has_lq = true;
        }
        next_rp = rp + 1;
    }
    if (has_lq) try_close(next_rp, n);
}

/* ================================================================= genotype rules */

namespace {

inline bool seq_eq(const CandSet &cs, uint32_t a, uint32_t b) {
    return cs.seq_len[a] == cs.seq_len[b] &&
           memcmp(cs.pool + cs.seq_off[a], cs.pool + cs.seq_off[b], cs.seq_len[a]) == 0;
}
inline size_t min_count_for(size_t c) { return c >= 9 ? 3 : (c >= 6 ? 2 : 1); }  // main.rs:803-811

struct Stat {
    size_t max1_c = 0, max1_p = 0, max2_c = 0, max2_p = 0;
    size_t per_pos[64];                                  // group size seen from the group's first scored member
    std::vector<std::pair<uint32_t, size_t>> by_order;  // order of a group's first scored member -> size
    size_t *find(uint32_t order) {
        for (auto &x : by_order)
            if (x.first == order) return &x.second;
        return nullptr;
    }
    size_t get(uint32_t order) {
        size_t *p = find(order);
        return p ? *p : 0;
    }
    void set(uint32_t order, size_t v) {
        size_t *p = find(order);
        if (p) *p = v;
        else by_order.emplace_back(order, v);
    }
};

// fill_order_stat (main.rs:813-849)
void order_stat(const CandSet &cs, const std::vector<uint32_t> &cand, Stat &st) {
    const size_t n = cand.size();
    st = Stat();
    std::fill(st.per_pos, st.per_pos + 64, 0);
    // first index holding the same string
    uint32_t rep[64];
    for (size_t i = 0; i < n; i++) {
        rep[i] = (uint32_t)i;
        for (size_t j = 0; j < i; j++)
            if (rep[j] == j && seq_eq(cs, cand[i], cand[j])) {
                rep[i] = (uint32_t)j;
                break;
            }
    }
    for (size_t p1 = 0; p1 < n; p1++) {
        if (cs.kscore[cand[p1]] == 0 || st.per_pos[p1] > 0) continue;
        size_t c = 0;
        for (size_t x = p1; x < n; x++) c += rep[x] == rep[p1];
        st.set(cs.order[cand[p1]], c);
        for (size_t x = p1; x < n; x++)
            if (rep[x] == rep[p1]) st.per_pos[x] = c;
        if (c > st.max1_c || (c == st.max1_c && cs.order[cand[p1]] == 0)) {
            st.max2_c = st.max1_c;
            st.max2_p = st.max1_p;
            st.max1_c = c;
            st.max1_p = p1;
        } else if (st.max1_p == st.max2_p || c > st.max2_c) {
            st.max2_c = c;
            st.max2_p = p1;
        }
    }
}

// is_valid_snp (main.rs:780-801): do the homopolymer-compressed strings differ?
bool hp_compressed_differ(const uint8_t *a, size_t na, const uint8_t *b, size_t nb) {
    size_t i = 0, j = 0;
    while (i < na && j < nb) {
        if (a[i] != b[j]) return true;
        while (i + 1 < na && a[i] == a[i + 1]) i++;
        while (j + 1 < nb && b[j] == b[j + 1]) j++;
        i++;
        j++;
    }
    return false;
}

}  // namespace

void mark_hete(CandSet &cs, std::vector<RegionState> &rs) {
    Stat st;
    for (auto &r : rs) {
        if (r.cand.empty()) continue;
        order_stat(cs, r.cand, st);
        const size_t min_c = min_count_for(r.cand.size());
        const uint32_t a = r.cand[st.max1_p], b = r.cand[st.max2_p];
        if (st.max2_c >= min_c &&
            (cs.seq_len[a] == cs.seq_len[b] || (r.cand.size() >= 6 && st.max2_c >= st.max1_c / 2)) &&
            hp_compressed_differ(cs.pool + cs.seq_off[a], cs.seq_len[a], cs.pool + cs.seq_off[b], cs.seq_len[b])) {
            r.lable |= LABLE_HETE;
            for (size_t p = 0; p < r.cand.size(); p++)
                if (cs.kscore[r.cand[p]] > 0 && st.per_pos[p] < min_c) cs.kscore[r.cand[p]] = 0;
        }
    }
}

void fill_seed(const CandSet &cs, std::vector<RegionState> &rs, long max_indel_len) {
    Stat st;
    for (auto &r : rs) {
        if (r.cand.empty()) herr(NP2_ERR_FORMAT, "LQ region without any candidate (reference would panic)");
        order_stat(cs, r.cand, st);
        auto str = [&](uint32_t c) { return std::string((const char *)cs.pool + cs.seq_off[c], cs.seq_len[c]); };
        r.sudoseed = str(r.cand[st.max1_p]);
        r.lable |= LABLE_SUCC | LABLE_RECH;
        const size_t min_c = min_count_for(r.cand.size());
        if (cs.order[r.cand[0]] != 0) herr(NP2_ERR_FORMAT, "the first lqseq is not ref.");
        // keep the reference allele when it has support (main.rs:876-890)
        if (size_t *v = st.find(0)) {
            if (*v > 1 && *v < min_c) *v = min_c;
        } else {
            size_t c = 0;
            for (uint32_t x : r.cand) c += seq_eq(cs, x, r.cand[0]);
            if (c > 1) st.set(0, min_c);
        }
        bool no_dup = true;  // no_dupseq_lqseq main.rs:851-860
        for (size_t p1 = 1; p1 < r.cand.size() && no_dup; p1++)
            for (size_t p2 = p1 + 1; p2 < r.cand.size(); p2++)
                if (seq_eq(cs, r.cand[p1], r.cand[p2])) {
                    no_dup = false;
                    break;
                }
        if (st.max1_p != 0 && st.max1_c < min_c && (st.max1_c > 1 || no_dup)) {
            st.set(cs.order[r.cand[st.max1_p]], min_c);
            st.set(0, min_c);
        } else if (st.max1_c < min_c) {
            st.set(0, min_c);
        }
        // retain_sort_seqs main.rs:714-726
        std::stable_sort(r.cand.begin(), r.cand.end(),
                         [&](uint32_t x, uint32_t y) { return st.get(cs.order[x]) > st.get(cs.order[y]); });
        size_t keep = 0;
        while (keep < r.cand.size() && st.get(cs.order[r.cand[keep]]) >= min_c) keep++;
        r.cand.resize(keep);
        if (r.cand.empty()) herr(NP2_ERR_FORMAT, "no candidate survives retain_sort_seqs (reference would panic)");
        const bool skip_long = std::labs((long)r.sudoseed.size() - (long)cs.seq_len[r.cand[0]]) > max_indel_len;
        if (r.cand.size() <= 1 || skip_long) {
            r.sudoseed = str(r.cand[0]);
            r.lable ^= LABLE_RECH;
            r.cand.clear();
        }
    }
}

/* ================================================================= phasing: graph + Louvain */

namespace {

typedef std::map<uint32_t, std::map<uint32_t, float>> Adj;  // ordered: ascending-id iteration (SURVEY hard part 3)

struct LNode {
    uint32_t id;
    float weight;
    std::vector<uint32_t> members;  // sorted original vertices
};

struct Level {
    Adj data;
    std::map<uint32_t, std::set<uint32_t>> comm;
    std::map<uint32_t, LNode> node;
};

void add_w(Adj &a, uint32_t x, uint32_t y, float w) {
    auto &m = a[x];
    auto it = m.find(y);
    if (it == m.end()) m.emplace(y, w);
    else it->second += w;
}

// louvain.rs:72-117
bool move_vertices(Level &lv) {
    bool moved_any = false;
    std::vector<std::pair<uint32_t, float>> acc;
    for (;;) {
        bool stop = true;
        for (auto &kv : lv.data) {
            const uint32_t v = kv.first;
            const uint32_t cur = lv.node.at(v).id;
            acc.clear();
            for (auto &e : kv.second) {
                const uint32_t c = lv.node.at(e.first).id;
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
                if (a.second > bw || (a.second == bw && a.first < bid)) {
                    bid = a.first;
                    bw = a.second;
                }
            if (bw > 0.0f && bid != cur) {
                lv.node.at(v).id = bid;
                lv.comm.at(bid).insert(v);
                lv.comm.at(cur).erase(v);
                stop = false;
                moved_any = true;
            }
        }
        if (stop) break;
    }
    return moved_any;
}

float internal_weight(const Level &lv, const std::set<uint32_t> &nodes) {
    float w = 0.f;
    for (uint32_t n : nodes) {
        w += lv.node.at(n).weight;
        auto it = lv.data.find(n);
        if (it != lv.data.end())
            for (auto &e : it->second)
                if (nodes.count(e.first)) w += e.second / 2.0f;
    }
    return w;
}

// louvain.rs:119-195
Level aggregate(Level &lv) {
    Level nx;
    std::vector<uint32_t> decluster;
    for (auto &kv : lv.comm) {
        if (kv.second.empty()) continue;
        LNode nn;
        nn.id = kv.first;
        nn.weight = internal_weight(lv, kv.second);
        for (uint32_t n : kv.second) {
            const auto &mm = lv.node.at(n).members;
            nn.members.insert(nn.members.end(), mm.begin(), mm.end());
        }
        std::sort(nn.members.begin(), nn.members.end());
        nn.members.erase(std::unique(nn.members.begin(), nn.members.end()), nn.members.end());
        if (nn.weight < 0.f) decluster.push_back(kv.first);
        else {
            nx.comm[kv.first] = {kv.first};
            nx.node[kv.first] = std::move(nn);
        }
    }
    for (uint32_t cid : decluster) {  // communities with negative internal weight fall apart again
        auto it = lv.comm.find(cid);
        if (it == lv.comm.end()) herr(NP2_ERR_FORMAT, "louvain: declustered community vanished (reference would panic)");
        std::set<uint32_t> nodes = std::move(it->second);
        lv.comm.erase(it);
        for (uint32_t nid : nodes) {
            uint32_t nn = nid;
            while (nx.comm.count(nn) || nx.node.count(nn)) nn++;
            nx.comm[nn] = {nn};
            LNode x;
            x.id = nn;
            x.weight = lv.node.at(nid).weight;
            x.members = lv.node.at(nid).members;
            nx.node[nn] = std::move(x);
            lv.comm[nn] = {nid};
        }
    }
    // edges between the (possibly re-keyed) communities: one pass over the old edges via member -> community
    std::unordered_map<uint32_t, uint32_t> owner;
    bool clean = true;
    for (auto &kv : lv.comm)
        for (uint32_t n : kv.second)
            if (!owner.emplace(n, kv.first).second) clean = false;
    if (clean) {
        Adj sum;
        for (auto &kv : lv.comm)
            for (uint32_t v : kv.second) {
                auto it = lv.data.find(v);
                if (it == lv.data.end()) continue;
                for (auto &e : it->second) {
                    auto ow = owner.find(e.first);
                    if (ow == owner.end() || !(ow->second > kv.first)) continue;
                    add_w(sum, kv.first, ow->second, e.second);
                }
            }
        for (auto &a : sum)
            for (auto &b : a.second)
                if (b.second != 0.f) {
                    add_w(nx.data, a.first, b.first, b.second);
                    add_w(nx.data, b.first, a.first, b.second);
                }
    } else {  // a vertex listed in two communities (reference quirk): literal pairwise form
        for (auto &c1 : lv.comm) {
            if (c1.second.empty()) continue;
            for (auto &c2 : lv.comm) {
                if (!(c2.first > c1.first) || c2.second.empty()) continue;
                float w = 0.f;
                for (uint32_t v : c1.second) {
                    auto it = lv.data.find(v);
                    if (it != lv.data.end())
                        for (auto &e : it->second)
                            if (c2.second.count(e.first)) w += e.second;
                }
                if (w != 0.f) {
                    add_w(nx.data, c1.first, c2.first, w);
                    add_w(nx.data, c2.first, c1.first, w);
                }
            }
        }
    }
    return nx;
}

struct Community {
    uint32_t id;
    float weight;
    std::vector<uint32_t> members;
};

}  // namespace

std::vector<uint32_t> phase_reads(const CandSet &cs, const std::vector<RegionState> &rs, bool asref,
                                  bool use_all_reads) {
    // ---- pair agreement over heterozygous regions (main.rs:953-992)
    struct Edge {
        uint32_t a, b;
        int32_t w;
    };
    std::vector<Edge> edges;
    std::map<uint32_t, float> ref_w;
    bool have_ref = false;
    std::set<uint32_t> invalid;
    for (auto &r : rs) {
        if (!(r.lable & LABLE_HETE)) continue;
        const size_t n = r.cand.size();
        uint32_t rep[64];
        for (size_t i = 0; i < n; i++) {
            rep[i] = (uint32_t)i;
            for (size_t j = 0; j < i; j++)
                if (rep[j] == j && seq_eq(cs, r.cand[i], r.cand[j])) {
                    rep[i] = (uint32_t)j;
                    break;
                }
        }
        for (size_t i = 0; i < n; i++) {
            if (cs.kscore[r.cand[i]] == 0) continue;
            const uint32_t oi = cs.order[r.cand[i]];
            for (size_t j = i + 1; j < n; j++) {
                if (cs.kscore[r.cand[j]] == 0) continue;
                const uint32_t oj = cs.order[r.cand[j]];
                const int32_t w = rep[i] == rep[j] ? 1 : -1;
                if (oi == 0) {
                    if (asref) {
                        ref_w[oj] += (float)w;
                        have_ref = true;
                    }
                    if (w < 0 && !use_all_reads) invalid.insert(oj);
                    continue;
                }
                if (oj == 0) herr(NP2_ERR_FORMAT, "seq2 order is equal to 0");
                edges.push_back({std::min(oi, oj), std::max(oi, oj), w});
            }
        }
    }
    std::sort(edges.begin(), edges.end(), [](const Edge &x, const Edge &y) { return x.a != y.a ? x.a < y.a : x.b < y.b; });
    Level lv;
    for (size_t i = 0; i < edges.size();) {
        size_t j = i;
        int32_t sum = 0, ndif = 0;
        while (j < edges.size() && edges[j].a == edges[i].a && edges[j].b == edges[i].b) {
            sum += edges[j].w;
            ndif += edges[j].w < 0 ? -1 : 0;
            j++;
        }
        const float w = ndif <= -3 ? (float)ndif : (float)sum;  // main.rs:996-1002
        lv.data[edges[i].a][edges[i].b] = w;
        lv.data[edges[i].b][edges[i].a] = w;
        i = j;
    }
    if (!use_all_reads) {  // main.rs:1004-1010
        for (uint32_t x : invalid) lv.data.erase(x);
        for (auto &kv : lv.data)
            for (uint32_t x : invalid) kv.second.erase(x);
    }
    // ---- Louvain (louvain.rs:59-257)
    for (auto &kv : lv.data) {
        lv.comm[kv.first] = {kv.first};
        lv.node[kv.first] = LNode{kv.first, 0.f, {kv.first}};
    }
    while (move_vertices(lv)) lv = aggregate(lv);
    std::vector<Community> comms;
    for (auto &kv : lv.comm) {  // get_communities louvain.rs:197-245
        if (kv.second.empty()) continue;
        Community c;
        c.id = kv.first;
        c.weight = internal_weight(lv, kv.second);
        for (uint32_t n : kv.second) {
            const auto &mm = lv.node.at(n).members;
            c.members.insert(c.members.end(), mm.begin(), mm.end());
        }
        comms.push_back(std::move(c));
    }
    Adj conflict;
    for (auto &c1 : comms)
        for (auto &c2 : comms) {
            if (!(c2.id > c1.id)) continue;
            float w = 0.f;
            for (uint32_t n1 : lv.comm.at(c1.id)) {
                auto it = lv.data.find(n1);
                if (it == lv.data.end()) continue;
                for (uint32_t n2 : lv.comm.at(c2.id)) {
                    auto jt = it->second.find(n2);
                    if (jt != it->second.end()) w += jt->second;
                }
            }
            if (w != 0.f) {
                if (!(w < 0.f)) herr(NP2_ERR_FORMAT, "the weight of two conflicting community is not less than 0");
                add_w(conflict, c1.id, c2.id, w);
                add_w(conflict, c2.id, c1.id, w);
            }
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

void splice(const Regions &rg, const std::vector<RegionState> &rs, uint8_t lable, const Cns &in, Cns &out) {
    out.pos.clear();
    out.base.clear();
    out.pos.reserve(in.pos.size());
    out.base.reserve(in.pos.size());
    const size_t nr = rs.size();
    // regions are stored in descending position; walk them from the last (lowest position) to the first
    auto next = [&](size_t i) {  // get_lqseqs_next_idx_by_lable main.rs:1017-1025
        i -= 1;
        while (i < nr && !(rs[i].lable & lable)) i -= 1;
        return i;
    };
    size_t ri = next(nr);
    const size_t n = in.pos.size();
    size_t i = 0;
    while (i < n) {
        const uint32_t p = in.pos[i];
        if (ri < nr && p == rg.start[ri]) {
            for (char b : rs[ri].sudoseed) {
                out.pos.push_back(p);
                out.base.push_back((uint8_t)b);
            }
            while (i < n && in.pos[i] <= rg.end[ri]) i++;
            ri = next(ri);
        } else {
            // copy the stretch up to the next region start in one go
            size_t j = i + 1;
            if (ri < nr) {
                const uint32_t stop = rg.start[ri];
                while (j < n && in.pos[j] != stop) j++;
            } else {
                j = n;
            }
            out.pos.insert(out.pos.end(), in.pos.begin() + i, in.pos.begin() + j);
            out.base.insert(out.base.end(), in.base.begin() + i, in.base.begin() + j);
            i = j;
        }
    }
}

namespace {
// first index with pos >= p / first index with pos > p (consensus positions are non-decreasing)
inline size_t lb(const Cns &c, uint32_t p) { return std::lower_bound(c.pos.begin(), c.pos.end(), p) - c.pos.begin(); }
inline size_t ub(const Cns &c, uint32_t p) { return std::upper_bound(c.pos.begin(), c.pos.end(), p) - c.pos.begin(); }

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
}  // namespace

void reupdate_build(const Regions &rg, const CandSet &cs, const std::vector<RegionState> &rs, const Cns &cns, uint32_t k,
                    Reupdate &ru) {
    ru = Reupdate();
    for (size_t i = rs.size(); i-- > 0;)
        if (rs[i].lable & LABLE_RECH) ru.rech.push_back((uint32_t)i);
    ru.off.push_back(0);
    const size_t N = cns.pos.size();
    auto put_cns = [&](size_t a, size_t b) {
        if (a < b) ru.pool.insert(ru.pool.end(), cns.base.begin() + a, cns.base.begin() + b);
    };
    auto put_cand = [&](uint32_t c) {
        ru.pool.insert(ru.pool.end(), cs.pool + cs.seq_off[c], cs.pool + cs.seq_off[c] + cs.seq_len[c]);
    };
    size_t sj = 0;
    while (sj < ru.rech.size()) {
        size_t ej = sj + 1;  // chain regions closer than k, at most 6 per group (main.rs:1197-1206)
        while (ej < ru.rech.size() && rg.start[ru.rech[ej]] < rg.end[ru.rech[ej - 1]] + k) {
            ej++;
            if (ej > sj + 5) break;
        }
        // flanks: k-1 consensus bases on either side (iter_consensus_extend main.rs:1100-1139)
        const size_t li = lb(cns, rg.start[ru.rech[sj]]);
        const size_t l0 = li > k - 1 ? li - (k - 1) : 0;
        const size_t rl = ub(cns, rg.end[ru.rech[ej - 1]]);  // index after the last base with pos <= end
        if (li >= N || rl == 0 || rl >= N + 1) herr(NP2_ERR_FORMAT, "consensus index out of range in reupdate");
        const size_t r1 = std::min(N, rl + (k - 1));
        Reupdate::Group g{(uint32_t)sj, (uint32_t)ej, ru.off.size() - 1};
        if (ej == sj + 1) {
            for (uint32_t c : rs[ru.rech[sj]].cand) {
                put_cns(l0, li);
                put_cand(c);
                put_cns(rl, r1);
                ru.off.push_back(ru.pool.size());
            }
        } else {
            std::vector<uint32_t> lens;
            for (size_t x = sj; x < ej; x++) lens.push_back((uint32_t)rs[ru.rech[x]].cand.size());
            // consensus between consecutive regions of the chain (iter_consensus_region main.rs:1068-1097)
            std::vector<std::pair<size_t, size_t>> mid;
            for (size_t x = sj; x + 1 < ej; x++) {
                const uint32_t s = rg.end[ru.rech[x]], e = rg.start[ru.rech[x + 1]];
                if (s + 1 == e) mid.emplace_back(0, 0);
                else mid.emplace_back(ub(cns, s), lb(cns, e));
            }
            for_each_choice(lens, [&](const std::vector<uint32_t> &ch) {
                put_cns(l0, li);
                for (size_t x = 0; x < ch.size(); x++) {
                    put_cand(rs[ru.rech[sj + x]].cand[ch[x]]);
                    if (x + 1 < ch.size()) put_cns(mid[x].first, mid[x].second);
                }
                put_cns(rl, r1);
                ru.off.push_back(ru.pool.size());
                if (ru.pool.size() > (1ull << 32)) herr(NP2_ERR_UNSUPPORTED, "cartesian re-check group too large");
            });
        }
        ru.groups.push_back(g);
        sj = ej;
    }
}

void reupdate_apply(const Regions &rg, CandSet &cs, std::vector<RegionState> &rs, const Reupdate &ru,
                    const uint16_t *ks, uint32_t iter_count, const Cns &in, Cns &out) {
    for (auto &g : ru.groups) {
        uint64_t si = g.first_string;
        if (g.ej == g.sj + 1) {
            for (uint32_t c : rs[ru.rech[g.sj]].cand) cs.kscore[c] = ks[si++];
        } else {
            std::vector<uint32_t> lens;
            for (size_t x = g.sj; x < g.ej; x++) lens.push_back((uint32_t)rs[ru.rech[x]].cand.size());
            for (size_t x = g.sj; x < g.ej; x++)
                for (uint32_t c : rs[ru.rech[x]].cand) cs.kscore[c] = 0;
            // later combinations overwrite earlier ones (main.rs:1351-1366)
            for_each_choice(lens, [&](const std::vector<uint32_t> &ch) {
                const uint16_t s = ks[si++];
                if (s > 0)
                    for (size_t x = 0; x < ch.size(); x++) cs.kscore[rs[ru.rech[g.sj + x]].cand[ch[x]]] = s;
            });
        }
    }
    // choose the allele (main.rs:1371-1406)
    for (auto &r : rs) {
        if (!(r.lable & LABLE_RECH)) continue;
        size_t c = 0, valid = 0;
        for (size_t p = 0; p < r.cand.size(); p++)
            if (cs.kscore[r.cand[p]] != 0) {
                if (c == 0 || cs.order[r.cand[p]] == 0) c = p + 1;
                valid++;
            }
        if (valid > 1) r.lable |= LABLE_TEMP;
        auto str = [&](uint32_t x) { return std::string((const char *)cs.pool + cs.seq_off[x], cs.seq_len[x]); };
        if (c != 0) {
            r.sudoseed = str(r.cand[c - 1]);
        } else if (iter_count == 1) {
            size_t i = 0;
            for (size_t p = 0; p < r.cand.size(); p++)
                if (cs.order[r.cand[p]] == 0) {
                    i = p;
                    break;
                }
            r.sudoseed = str(r.cand[i]);
        }
    }
    splice(rg, rs, LABLE_RECH, in, out);
    for (auto &r : rs) {  // main.rs:1411-1417
        if (!(r.lable & LABLE_RECH)) continue;
        if (r.lable & LABLE_TEMP) r.lable ^= LABLE_TEMP;
        else r.lable ^= LABLE_RECH;
    }
}

}  // namespace np2
