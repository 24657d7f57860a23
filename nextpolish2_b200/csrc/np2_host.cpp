// np2_host.cpp — host phases of the polish path (see np2_host.h).  Citations are to the reference
// (Nextomics/NextPolish2 @ 283dc5a).  Nothing here calls into oracle/.
#include "np2_host.h"

#include "np2_error.h"

#include <algorithm>
#include <cstring>
#include <map>
#include <set>
#include <stdexcept>
#include <thread>
#include <unordered_map>

namespace np2 {

[[noreturn]] static void herr(int code, const std::string &m) { throw Error(code, m); }

/* ================================================================= ingest */

namespace {
typedef Ingest::RecOut RecOut;
struct ParseErr {
    int64_t rec = INT64_MAX;
    std::string msg;
};
}  // namespace

// Record-level filter (main.rs:1758-1771) and fill_with_cigar bookkeeping (main.rs:386-440) without materialising
// the gapped strings.  Pass 1 walks the block_size chain (sequential by nature); pass 2 processes the records'
// CIGARs on all host threads; pass 3 concatenates.  The first failing record in file order decides the error,
// like the reference's sequential loop would.
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
    op_off.clear();
    op_col.clear();
    op_q.clear();
    op_t.clear();
    op_cig.clear();
    nib_off.clear();
    ck_off.clear();
    total_cols = 0;
    rec_off.clear();
    ro.clear();
}

void parse_records(const uint8_t *bam, uint64_t bam_len, uint32_t tlen, const np2_opts &opt, Ingest &out) {
    out.clear();
    std::vector<uint64_t> &rec_off = out.rec_off;
    {
        uint64_t off = 0;
        while (off + 4 <= bam_len) {
            int32_t bs;
            memcpy(&bs, bam + off, 4);
            if (bs < 32 || off + 4 + (uint64_t)bs > bam_len) herr(NP2_ERR_FORMAT, "BAM/SAM parsing failed!");
            rec_off.push_back(off + 4);
            off += 4 + (uint64_t)bs;
        }
        if (off != bam_len) herr(NP2_ERR_FORMAT, "BAM/SAM parsing failed!");
    }
    const size_t nrec = rec_off.size();
    out.all_tid.resize(nrec);
    out.all_pos.resize(nrec);
    std::vector<RecOut> &ro = out.ro;
    ro.assign(nrec, RecOut());
    unsigned T = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    if (nrec < 2048) T = 1;
    auto &t_col = out.t_col, &t_q = out.t_q, &t_t = out.t_t, &t_cig = out.t_cig;
    if (t_col.size() < T) {
        t_col.resize(T);
        t_q.resize(T);
        t_t.resize(T);
        t_cig.resize(T);
    }
    std::vector<ParseErr> t_err(T);
    auto work = [&](unsigned ti) {
        const size_t b = nrec * ti / T, e = nrec * (ti + 1) / T;
        // thread-local vectors: the shared arrays of vector headers would false-share on every push_back
        std::vector<uint32_t> vcol, vq, vt, vcig;
        vcol.swap(t_col[ti]);  // take last job's capacity
        vq.swap(t_q[ti]);
        vt.swap(t_t[ti]);
        vcig.swap(t_cig[ti]);
        vcol.clear();
        vq.clear();
        vt.clear();
        vcig.clear();
        const size_t guess = (e - b) * 64;
        vcol.reserve(guess);
        vq.reserve(guess);
        vt.reserve(guess);
        vcig.reserve(guess);
        struct Publish {
            std::vector<uint32_t> &a, &b, &c, &d, &A, &B, &C, &D;
            ~Publish() {
                A.swap(a);
                B.swap(b);
                C.swap(c);
                D.swap(d);
            }
        } publish{vcol, vq, vt, vcig, t_col[ti], t_q[ti], t_t[ti], t_cig[ti]};
        ParseErr my_err;
        struct PublishErr {
            ParseErr &src, &dst;
            ~PublishErr() { dst = src; }
        } publish_err{my_err, t_err[ti]};
        auto fail = [&](size_t rec, const char *m) {
            if ((int64_t)rec < my_err.rec) {
                my_err.rec = (int64_t)rec;
                my_err.msg = m;
            }
        };
        for (size_t rec = b; rec < e; rec++) {
            const uint8_t *r = bam + rec_off[rec];
            int32_t bs, ref_id, pos, l_seq;
            uint16_t n_cig, flag;
            memcpy(&bs, r - 4, 4);
            memcpy(&ref_id, r, 4);
            memcpy(&pos, r + 4, 4);
            const uint32_t l_name = r[8], mapq = r[9];
            memcpy(&n_cig, r + 12, 2);
            memcpy(&flag, r + 14, 2);
            memcpy(&l_seq, r + 16, 4);
            out.all_tid[rec] = ref_id;
            out.all_pos[rec] = pos;
            if (l_seq < 0 || 32ull + l_name + 4ull * n_cig + ((uint64_t)l_seq + 1) / 2 > (uint64_t)bs) {
                fail(rec, "BAM/SAM parsing failed!");
                return;  // a sequential reader stops here
            }
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
            if (pos < 0 || (uint64_t)pos > tlen) {
                fail(rec, "alignment starts outside the contig");
                return;
            }
            uint32_t qs = 0, ts = 0, col = 0, aln_q_s = 0, aln_q_e = 0, n_ops = 0;
            bool first = true;
            const char *bad = nullptr;
            for (uint32_t i = 0; i < n_cig && !bad; i++) {
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
                        if (op != 2 && (uint64_t)qs + l > (uint64_t)l_seq) {
                            bad = "CIGAR consumes more query bases than SEQ holds";
                            break;
                        }
                        if (op != 1 && (uint64_t)pos + ts + l > tlen) {
                            bad = "alignment runs past the end of the contig";
                            break;
                        }
                        if (l) {
                            vcol.push_back(col);
                            vq.push_back(qs);
                            vt.push_back(ts);
                            vcig.push_back(c);
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
            if (bad) {
                fail(rec, bad);
                return;
            }
            if (aln_q_e == 0) aln_q_e = qs;
            RecOut &o = ro[rec];
            o.kept = 1;
            o.is_clip = (uint32_t)(aln_q_e - aln_q_s + opt.max_clip_len) < (uint32_t)rlen ? 1 : 0;  // main.rs:1796
            o.ncols = col;
            o.rlen = (uint32_t)rlen;
            o.rspan = (uint32_t)rspan;
            o.n_ops = n_ops;
            o.seq_off = rec_off[rec] + 32 + l_name + 4ull * n_cig;
        }
    };
    if (T == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (unsigned ti = 0; ti < T; ti++) th.emplace_back(work, ti);
        for (auto &t : th) t.join();
    }
    const ParseErr *first_err = nullptr;
    for (auto &e : t_err)
        if (e.rec != INT64_MAX && (!first_err || e.rec < first_err->rec)) first_err = &e;
    if (first_err) herr(NP2_ERR_FORMAT, first_err->msg);
    // pass 3: compact the kept records, concatenate the per-thread op lists (already in record order)
    size_t nk = 0, nops = 0;
    for (auto &o : ro) nk += o.kept, nops += o.n_ops;
    out.rec_idx.reserve(nk);
    out.pos.reserve(nk);
    out.ncols.reserve(nk);
    out.rlen.reserve(nk);
    out.rspan.reserve(nk);
    out.is_clip.reserve(nk);
    out.seq_off.reserve(nk);
    out.op_off.reserve(nk + 1);
    out.nib_off.reserve(nk + 1);
    out.ck_off.reserve(nk + 1);
    out.op_off.push_back(0);
    out.nib_off.push_back(0);
    out.ck_off.push_back(0);
    for (size_t rec = 0; rec < nrec; rec++) {
        const RecOut &o = ro[rec];
        if (!o.kept) continue;
        out.rec_idx.push_back((int32_t)rec);
        out.pos.push_back((uint32_t)out.all_pos[rec]);
        out.ncols.push_back(o.ncols);
        out.rlen.push_back(o.rlen);
        out.rspan.push_back(o.rspan);
        out.is_clip.push_back(o.is_clip);
        out.seq_off.push_back(o.seq_off);
        out.op_off.push_back(out.op_off.back() + o.n_ops);
        out.nib_off.push_back(out.nib_off.back() + ((((uint64_t)o.ncols / 16 + 1) * 8 + 15) & ~15ull));
        out.ck_off.push_back(out.ck_off.back() + (o.ncols + 31) / 32);
        out.total_cols += o.ncols;
    }
    out.op_col.resize(nops);
    out.op_q.resize(nops);
    out.op_t.resize(nops);
    out.op_cig.resize(nops);
    size_t w = 0;
    for (unsigned ti = 0; ti < T; ti++) {
        const size_t n = t_col[ti].size();
        if (n) {
            memcpy(out.op_col.data() + w, t_col[ti].data(), n * 4);
            memcpy(out.op_q.data() + w, t_q[ti].data(), n * 4);
            memcpy(out.op_t.data() + w, t_t[ti].data(), n * 4);
            memcpy(out.op_cig.data() + w, t_cig[ti].data(), n * 4);
        }
        w += n;
    }
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

std::vector<uint32_t> phase_reads(const uint64_t *keys, const long long *vals, uint64_t n_edges, bool asref,
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
