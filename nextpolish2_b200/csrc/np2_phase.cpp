// np2_phase.cpp — the host half of phase_reads_by_lqseqs (main.rs:994-1015) + louvain.rs:59-356, flat arrays.
//
// Input: the read x read agreement edges reduced on the device (one record per read pair).  Everything here is
// CSR / id-indexed arrays: no node-based containers on the common path.  Why that is exact:
//   * all weights are small integers or halves of them (sums of +-1, "/ 2.0" of such sums), so every f32 sum is
//     exact and independent of the order of accumulation; only the ORDER OF DECISIONS matters, and that is kept:
//     vertices are visited in ascending id (louvain.rs:77), ties go to the smaller community id (99-101), a move
//     needs a positive weight and a different community (103);
//   * a level's communities are the groups of equal `cid`; the reference's `communities` map holds exactly those
//     sets (plus empty ones it skips), so they are rebuilt from `cid` after the moves instead of being maintained.
// The one thing not handled here is a community whose internal weight is negative (louvain.rs:136-165: it falls
// apart again and its vertices are re-keyed, with quirks): when one shows up the call is served from scratch by the
// general map/set implementation (phase_reads_general in np2_host.cpp), which reproduces those quirks.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#include "../../include/np2gpu.h"
#include "np2_error.h"
#include "np2_host.h"

namespace np2 {

namespace {

thread_local int g_phase_path = 0;
thread_local float g_phase_ms[4] = {0, 0, 0, 0};  // build, move, aggregate, communities

struct Lvl {
    uint32_t n = 0;                   // every array below is indexed by vertex / community id < n
    std::vector<uint32_t> verts;      // all vertices of the level, ascending (keys of `communities`)
    std::vector<uint32_t> ids;        // those that appear in `data` (louvain.rs:77 iterates these), ascending
    // adjacency (CSR), neighbours ascending: views, so that level 0 can sit in the buffer it was downloaded into
    const uint32_t *aoff = nullptr, *ato = nullptr;
    const float *aw = nullptr;
    std::vector<uint32_t> aoff_own, ato_own;
    std::vector<float> aw_own;
    std::vector<uint32_t> cid;        // vertex -> community
    std::vector<float> nweight;       // Node.weight
    std::vector<uint32_t> moff, mdat; // Node.nodes: original vertices of each vertex (CSR)
    void bind_own() {
        aoff = aoff_own.data();
        ato = ato_own.data();
        aw = aw_own.data();
    }
};

// louvain.rs:72-117.  The outcome of visiting a vertex depends only on its neighbours' communities, so a vertex is
// re-examined only when one of them changed since its last visit (a skipped visit would have moved nothing).
// Works on ids[i0, i1); slot / stamp / dirty are indexed by vertex id and shared between concurrent calls, which touch
// disjoint id ranges (see move_vertices).
bool move_range(Lvl &lv, size_t i0, size_t i1, uint32_t *slot, uint32_t *stamp, uint8_t *dirty) {
    bool moved_any = false;
    // per-visit accumulator: weight towards each neighbouring community (the choice below only depends on the sums, not
    // on the order the communities were met in).  A read's neighbours sit in a handful of communities, so the first
    // kSmall of them live in a small array that is searched linearly (registers / L1); only a vertex that touches
    // more falls back to the stamped slot table.
    constexpr int kSmall = 8;
    uint32_t sc[kSmall];
    float sw[kSmall];
    std::vector<std::pair<uint32_t, float>> acc;
    uint32_t visit = 0;
    const uint32_t *aoff = lv.aoff, *ato = lv.ato;
    const float *aw = lv.aw;
    uint32_t *cid = lv.cid.data();
    for (;;) {
        bool stop = true;
        for (size_t i = i0; i < i1; i++) {
            const uint32_t v = lv.ids[i];
            if (!dirty[v]) continue;
            dirty[v] = 0;
            const uint32_t cur = cid[v];
            const uint32_t e0 = aoff[v], e1 = aoff[v + 1];
            if (e0 == e1) continue;
            int ns = 0;
            bool spilled = false;
            uint32_t e = e0;
            for (; e < e1; e++) {
                const uint32_t c = cid[ato[e]];
                int k = 0;
                while (k < ns && sc[k] != c) k++;
                if (k < ns) {
                    sw[k] += aw[e];
                } else if (ns < kSmall) {
                    sc[ns] = c;
                    sw[ns++] = aw[e];
                } else {
                    spilled = true;
                    break;
                }
            }
            uint32_t bid;
            float bw;
            if (!spilled) {
                bid = sc[0];
                bw = sw[0];
                for (int k = 1; k < ns; k++)
                    if (sw[k] > bw || (sw[k] == bw && sc[k] < bid)) {  // max weight, ties -> smaller id
                        bid = sc[k];
                        bw = sw[k];
                    }
            } else {  // many neighbouring communities: start over with the slot table
                acc.clear();
                if (++visit == 0) {  // stamp wrap-around (communities of this range are vertices of this range)
                    for (size_t x = i0; x < i1; x++) stamp[lv.ids[x]] = 0;
                    visit = 1;
                }
                for (e = e0; e < e1; e++) {
                    const uint32_t c = cid[ato[e]];
                    if (stamp[c] != visit) {
                        stamp[c] = visit;
                        slot[c] = (uint32_t)acc.size();
                        acc.emplace_back(c, aw[e]);
                    } else {
                        acc[slot[c]].second += aw[e];
                    }
                }
                bid = acc[0].first;
                bw = acc[0].second;
                for (auto &a : acc)
                    if (a.second > bw || (a.second == bw && a.first < bid)) {
                        bid = a.first;
                        bw = a.second;
                    }
            }
            if (bw > 0.0f && bid != cur) {
                cid[v] = bid;
                for (e = e0; e < e1; e++) dirty[ato[e]] = 1;
                stop = false;
                moved_any = true;
            }
        }
        if (stop) break;
    }
    return moved_any;
}
// The sweeps over all vertices (ascending id, repeated until nothing moves) split exactly into the same sweeps over
// every set of vertices that no edge leaves: a visit only reads and writes communities of the vertex's own neighbours,
// and a set in which a sweep moved nothing has no dirty vertex left, so later sweeps of the whole graph pass over
// it.  Reads are numbered along the contig and only overlapping reads share an edge, so such sets are found as id
// intervals (a cut wherever no earlier vertex has a neighbour at or beyond the next one: phase-block boundaries) and
// are moved concurrently on the host pool.
// adjacency entries below which a level is handled by one thread (NP2_PHASE_PAR_MIN lowers it so that small test graphs
// reach the concurrent paths)
uint64_t par_min_edges() {
    static const uint64_t v = [] {
        const char *e = getenv("NP2_PHASE_PAR_MIN");
        return e ? (uint64_t)strtoull(e, nullptr, 10) : 100000ull;
    }();
    return v;
}
// NP2_PHASE_CHECK=1: verify the invariants the concurrent paths rest on (neighbour lists ascending, no self loops)
void check_level(const Lvl &lv) {
    static const bool on = [] {
        const char *e = getenv("NP2_PHASE_CHECK");
        return e && atoi(e) != 0;
    }();
    if (!on) return;
    for (uint32_t v : lv.ids)
        for (uint32_t e = lv.aoff[v]; e < lv.aoff[v + 1]; e++) {
            if (lv.ato[e] == v) throw Error(NP2_ERR_INTERNAL, "phase graph: self loop");
            if (e > lv.aoff[v] && lv.ato[e - 1] >= lv.ato[e]) throw Error(NP2_ERR_INTERNAL, "phase graph: neighbour list not ascending");
        }
}
bool move_vertices(Lvl &lv) {
    std::vector<uint32_t> slot(lv.n, 0), stamp(lv.n, 0);
    std::vector<uint8_t> dirty(lv.n, 1);
    const size_t ni = lv.ids.size();
    const unsigned threads = host_threads();
    const uint64_t n_edges = ni ? lv.aoff[lv.ids[ni - 1] + 1] - lv.aoff[lv.ids[0]] : 0;
    check_level(lv);
    if (threads < 2 || n_edges < par_min_edges()) return move_range(lv, 0, ni, slot.data(), stamp.data(), dirty.data());
    // independent intervals, then chunks of whole intervals with about the same number of edges
    const uint64_t per_chunk = n_edges / (4 * threads) + 1;
    std::vector<size_t> chunk_begin(1, 0);
    uint64_t reach = 0, in_chunk = 0;  // reach: 1 + the largest neighbour id seen so far
    for (size_t i = 0; i < ni; i++) {
        const uint32_t v = lv.ids[i];
        if (i && reach <= v && in_chunk >= per_chunk) {
            chunk_begin.push_back(i);
            in_chunk = 0;
        }
        const uint32_t deg = lv.aoff[v + 1] - lv.aoff[v];
        if (deg) reach = std::max<uint64_t>(reach, (uint64_t)lv.ato[lv.aoff[v + 1] - 1] + 1);  // neighbours ascending
        reach = std::max<uint64_t>(reach, (uint64_t)v + 1);
        in_chunk += deg;
    }
    chunk_begin.push_back(ni);
    const size_t n_chunks = chunk_begin.size() - 1;
    if (n_chunks < 2) return move_range(lv, 0, ni, slot.data(), stamp.data(), dirty.data());
    std::atomic<size_t> next(0);
    std::atomic<int> moved(0);
    parallel_for((unsigned)std::min<size_t>(threads, n_chunks), [&](unsigned) {
        for (;;) {
            const size_t c = next.fetch_add(1);
            if (c >= n_chunks) break;
            if (move_range(lv, chunk_begin[c], chunk_begin[c + 1], slot.data(), stamp.data(), dirty.data())) moved.store(1);
        }
    });
    return moved.load() != 0;
}

// vertices grouped by community: comms ascending, members of each ascending
struct Groups {
    std::vector<uint32_t> comms, off, dat;  // off is indexed by position in comms
};
void group_by_cid(const Lvl &lv, Groups &g, std::vector<uint32_t> &slot) {
    slot.assign(lv.n, 0);
    for (uint32_t v : lv.verts) slot[lv.cid[v]]++;
    g.comms.clear();
    g.off.assign(1, 0);
    for (uint32_t c = 0; c < lv.n; c++)
        if (slot[c]) {
            g.comms.push_back(c);
            g.off.push_back(g.off.back() + slot[c]);
            slot[c] = (uint32_t)g.comms.size() - 1;  // community -> its position
        }
    g.dat.resize(lv.verts.size());
    std::vector<uint32_t> cur(g.off.begin(), g.off.end() - 1);
    for (uint32_t v : lv.verts) g.dat[cur[slot[lv.cid[v]]]++] = v;
}
// Node.weight of a community: its vertices' weights + every internal edge once (seen from both ends, halved)
float internal_weight(const Lvl &lv, const Groups &g, size_t gi) {
    const uint32_t c = g.comms[gi];
    float w = 0.f;
    for (uint32_t x = g.off[gi]; x < g.off[gi + 1]; x++) {
        const uint32_t v = g.dat[x];
        w += lv.nweight[v];
        for (uint32_t e = lv.aoff[v]; e < lv.aoff[v + 1]; e++)
            if (lv.cid[lv.ato[e]] == c) w += lv.aw[e] / 2.0f;
    }
    return w;
}
// sum of the edges between every pair of communities (c1 < c2), emitted in ascending (c1, c2)
struct PairSum {
    uint32_t c1, c2;
    float w;
};
void between_range(const Lvl &lv, const Groups &g, size_t g0, size_t g1, std::vector<PairSum> &out) {
    std::vector<uint32_t> stamp(lv.n, 0), touched;
    std::vector<float> acc(lv.n, 0.f);
    for (size_t gi = g0; gi < g1; gi++) {
        const uint32_t c = g.comms[gi];
        touched.clear();
        for (uint32_t x = g.off[gi]; x < g.off[gi + 1]; x++) {
            const uint32_t v = g.dat[x];
            for (uint32_t e = lv.aoff[v]; e < lv.aoff[v + 1]; e++) {
                const uint32_t o = lv.cid[lv.ato[e]];
                if (!(o > c)) continue;
                if (stamp[o] != c + 1) {
                    stamp[o] = c + 1;
                    acc[o] = 0.f;
                    touched.push_back(o);
                }
                acc[o] += lv.aw[e];
            }
        }
        std::sort(touched.begin(), touched.end());
        for (uint32_t o : touched) out.push_back({c, o, acc[o]});
    }
}
// chunks of whole communities with about the same number of member vertices (the sums of a community only read)
std::vector<size_t> community_chunks(const Lvl &lv, const Groups &g, unsigned &threads) {
    threads = host_threads();
    const uint64_t n_edges = lv.verts.empty() ? 0 : lv.aoff[lv.verts.back() + 1] - lv.aoff[lv.verts.front()];
    std::vector<size_t> cb(1, 0);
    if (threads < 2 || n_edges < par_min_edges() || g.comms.size() < 2) {
        cb.push_back(g.comms.size());
        threads = 1;
        return cb;
    }
    const uint64_t per = g.dat.size() / (4 * threads) + 1;
    uint64_t in_chunk = 0;
    for (size_t gi = 0; gi < g.comms.size(); gi++) {
        if (gi && in_chunk >= per) {
            cb.push_back(gi);
            in_chunk = 0;
        }
        in_chunk += g.off[gi + 1] - g.off[gi];
    }
    cb.push_back(g.comms.size());
    return cb;
}
void between_communities(const Lvl &lv, const Groups &g, std::vector<PairSum> &out) {
    out.clear();
    unsigned threads;
    const std::vector<size_t> cb = community_chunks(lv, g, threads);
    const size_t nc = cb.size() - 1;
    if (nc < 2) {
        between_range(lv, g, 0, g.comms.size(), out);
        return;
    }
    std::vector<std::vector<PairSum>> part(nc);
    std::atomic<size_t> next(0);
    parallel_for((unsigned)std::min<size_t>(threads, nc), [&](unsigned) {
        for (;;) {
            const size_t c = next.fetch_add(1);
            if (c >= nc) break;
            between_range(lv, g, cb[c], cb[c + 1], part[c]);
        }
    });
    for (auto &p : part) out.insert(out.end(), p.begin(), p.end());  // chunks in community order: ascending (c1, c2)
}
// Node.weight of every community, by position in g.comms
void internal_weights(const Lvl &lv, const Groups &g, std::vector<float> &w) {
    w.assign(g.comms.size(), 0.f);
    unsigned threads;
    const std::vector<size_t> cb = community_chunks(lv, g, threads);
    const size_t nc = cb.size() - 1;
    if (nc < 2) {
        for (size_t gi = 0; gi < g.comms.size(); gi++) w[gi] = internal_weight(lv, g, gi);
        return;
    }
    std::atomic<size_t> next(0);
    parallel_for((unsigned)std::min<size_t>(threads, nc), [&](unsigned) {
        for (;;) {
            const size_t c = next.fetch_add(1);
            if (c >= nc) break;
            for (size_t gi = cb[c]; gi < cb[c + 1]; gi++) w[gi] = internal_weight(lv, g, gi);
        }
    });
}

// louvain.rs:119-195 when no community has to be declustered; false when one has
bool aggregate(const Lvl &lv, Lvl &nx) {
    Groups g;
    std::vector<uint32_t> slot;
    group_by_cid(lv, g, slot);
    nx = Lvl();
    nx.n = lv.n;
    nx.cid.assign(lv.n, 0);
    nx.nweight.assign(lv.n, 0.f);
    nx.moff.assign(lv.n + 1, 0);
    nx.verts = g.comms;
    std::vector<float> iw;
    internal_weights(lv, g, iw);
    for (size_t gi = 0; gi < g.comms.size(); gi++) {
        const float w = iw[gi];
        if (w < 0.f) return false;
        const uint32_t c = g.comms[gi];
        nx.cid[c] = c;
        nx.nweight[c] = w;
        uint32_t m = 0;
        for (uint32_t x = g.off[gi]; x < g.off[gi + 1]; x++) m += lv.moff[g.dat[x] + 1] - lv.moff[g.dat[x]];
        nx.moff[c + 1] = m;
    }
    for (uint32_t c = 0; c < lv.n; c++) nx.moff[c + 1] += nx.moff[c];
    nx.mdat.resize(nx.moff[lv.n]);
    for (size_t gi = 0; gi < g.comms.size(); gi++) {
        const uint32_t c = g.comms[gi];
        uint32_t w = nx.moff[c];
        for (uint32_t x = g.off[gi]; x < g.off[gi + 1]; x++) {
            const uint32_t v = g.dat[x];
            for (uint32_t y = lv.moff[v]; y < lv.moff[v + 1]; y++) nx.mdat[w++] = lv.mdat[y];
        }
        std::sort(nx.mdat.begin() + nx.moff[c], nx.mdat.begin() + w);  // members of distinct vertices are disjoint
    }
    std::vector<PairSum> ps;
    between_communities(lv, g, ps);
    nx.aoff_own.assign(lv.n + 1, 0);
    for (auto &p : ps)
        if (p.w != 0.f) {
            nx.aoff_own[p.c1 + 1]++;
            nx.aoff_own[p.c2 + 1]++;
        }
    for (uint32_t c = 0; c < lv.n; c++) nx.aoff_own[c + 1] += nx.aoff_own[c];
    nx.ato_own.resize(nx.aoff_own[lv.n]);
    nx.aw_own.resize(nx.aoff_own[lv.n]);
    std::vector<uint32_t> cur(nx.aoff_own.begin(), nx.aoff_own.end() - 1);
    for (auto &p : ps)  // ascending (c1, c2): every list receives its smaller neighbours first, each part ascending
        if (p.w != 0.f) {
            nx.ato_own[cur[p.c1]] = p.c2;
            nx.aw_own[cur[p.c1]++] = p.w;
            nx.ato_own[cur[p.c2]] = p.c1;
            nx.aw_own[cur[p.c2]++] = p.w;
        }
    nx.bind_own();
    for (uint32_t c : nx.verts)
        if (nx.aoff[c + 1] > nx.aoff[c]) nx.ids.push_back(c);  // louvain.rs:183-186
    return true;
}

struct Community {
    uint32_t id;
    float weight;
    uint32_t gi;
};

}  // namespace

int phase_last_path() { return g_phase_path; }
const float *phase_last_ms() { return g_phase_ms; }

namespace {
// Louvain + phase_communities on a prepared level 0.  `general` serves the call from scratch when a community has to
// be declustered.
std::vector<uint32_t> phase_core(Lvl &lv, uint32_t n, const uint8_t *bad_v, const uint8_t *in_ref, const float *ref_w,
                                 bool have_ref, const std::function<std::vector<uint32_t>()> &general,
                                 std::chrono::steady_clock::time_point T0) {
    auto lap = [&](int slot) {
        auto t = std::chrono::steady_clock::now();
        g_phase_ms[slot] += std::chrono::duration<float, std::milli>(t - T0).count();
        T0 = t;
    };
    lap(0);
    // ---- Louvain (louvain.rs:59-257)
    for (;;) {
        const bool moved = move_vertices(lv);
        lap(1);
        if (!moved) break;
        Lvl nx;
        if (!aggregate(lv, nx)) {
            g_phase_path = 2;
            return general();
        }
        lv = std::move(nx);
        if (lv.aoff == nullptr || lv.aoff_own.data() != lv.aoff) lv.bind_own();  // views follow the moved vectors
        lap(2);
    }
    // get_communities (louvain.rs:197-245)
    Groups g;
    std::vector<uint32_t> slot;
    group_by_cid(lv, g, slot);
    std::vector<Community> comms;
    std::vector<float> iw;
    internal_weights(lv, g, iw);
    for (size_t gi = 0; gi < g.comms.size(); gi++) comms.push_back({g.comms[gi], iw[gi], (uint32_t)gi});
    std::vector<PairSum> ps;
    between_communities(lv, g, ps);
    std::vector<std::vector<uint32_t>> conflict(g.comms.size());  // by position in g.comms
    for (auto &p : ps) {
        if (p.w == 0.f) continue;
        if (!(p.w < 0.f)) throw Error(NP2_ERR_FORMAT, "the weight of two conflicting community is not less than 0");
        conflict[slot[p.c1]].push_back(slot[p.c2]);
        conflict[slot[p.c2]].push_back(slot[p.c1]);
    }
    // ---- phase_communities (louvain.rs:290-356)
    if (have_ref) {
        std::vector<std::pair<std::pair<int32_t, float>, size_t>> key;
        for (size_t i = 0; i < comms.size(); i++) {
            int32_t cnt = 0;
            float w = 0.f;
            const uint32_t gi = comms[i].gi;
            for (uint32_t x = g.off[gi]; x < g.off[gi + 1]; x++) {
                const uint32_t v = g.dat[x];
                for (uint32_t y = lv.moff[v]; y < lv.moff[v + 1]; y++) {
                    const uint32_t m = lv.mdat[y];
                    if (!in_ref[m]) continue;
                    if (ref_w[m] > 0.f) cnt++;
                    else if (ref_w[m] < 0.f) cnt--;
                    w += ref_w[m];
                }
            }
            key.push_back({{cnt, w}, i});
        }
        std::stable_sort(key.begin(), key.end(),
                         [](const std::pair<std::pair<int32_t, float>, size_t> &x,
                            const std::pair<std::pair<int32_t, float>, size_t> &y) { return x.first > y.first; });
        std::vector<Community> sorted;
        for (auto &k : key) sorted.push_back(comms[k.second]);
        comms.swap(sorted);
    } else {
        std::stable_sort(comms.begin(), comms.end(),
                         [](const Community &x, const Community &y) { return x.weight > y.weight; });
    }
    // keep the first, drop whatever conflicts with a kept one (louvain.rs:327-353)
    std::vector<uint8_t> bad(g.comms.size(), 0);
    std::vector<uint32_t> rank(g.comms.size(), 0);
    for (size_t p = 0; p < comms.size(); p++) rank[comms[p].gi] = (uint32_t)p;
    for (size_t p = 0; p < comms.size(); p++) {
        if (bad[comms[p].gi]) continue;
        for (uint32_t q : conflict[comms[p].gi])
            if (rank[q] > p) bad[q] = 1;
    }
    std::vector<uint32_t> out;
    for (uint32_t v = 0; v < n; v++)
        if (bad_v[v]) out.push_back(v);
    for (size_t gi = 0; gi < g.comms.size(); gi++) {
        if (!bad[gi]) continue;
        for (uint32_t x = g.off[gi]; x < g.off[gi + 1]; x++) {
            const uint32_t v = g.dat[x];
            out.insert(out.end(), lv.mdat.begin() + lv.moff[v], lv.mdat.begin() + lv.moff[v + 1]);
        }
    }
    std::sort(out.begin(), out.end());
    out.erase(std::unique(out.begin(), out.end()), out.end());
    lap(3);
    return out;
}
// level-0 vertices: every read that still has a pair (its members are itself)
void level0_vertices(Lvl &lv, uint32_t n, const uint8_t *has) {
    lv.n = n;
    lv.cid.resize(n);
    lv.nweight.assign(n, 0.f);
    lv.moff.assign(n + 1, 0);
    for (uint32_t v = 0; v < n; v++) {
        lv.cid[v] = v;
        lv.moff[v + 1] = lv.moff[v] + (has[v] ? 1 : 0);
        if (has[v]) {
            lv.verts.push_back(v);
            lv.mdat.push_back(v);
        }
    }
    lv.ids = lv.verts;
}
}  // namespace

// Level 0 as built on the device (np2_geno.cu k_phase_*): adjacency in CSR form with the `dif <= -3` override applied,
// invalid reads already removed, per-read flags.  The arrays are used in place.
std::vector<uint32_t> phase_reads_csr(uint32_t n, const uint32_t *aoff, const uint32_t *ato, const float *aw,
                                      const uint8_t *has, const uint8_t *bad_v, const uint8_t *in_ref, const float *ref_w,
                                      bool asref, const std::function<std::vector<uint32_t>()> &general) {
    g_phase_path = 1;
    for (float &x : g_phase_ms) x = 0.f;
    auto T0 = std::chrono::steady_clock::now();
    bool have_ref = false;
    if (asref)
        for (uint32_t v = 0; v < n && !have_ref; v++) have_ref = in_ref[v] != 0;
    Lvl lv;
    lv.aoff = aoff;
    lv.ato = ato;
    lv.aw = aw;
    level0_vertices(lv, n, has);
    return phase_core(lv, n, bad_v, in_ref, ref_w, have_ref, general, T0);
}

std::vector<uint32_t> phase_reads(const uint64_t *keys, const long long *vals, uint64_t n_edges, bool asref,
                                  bool use_all_reads) {
    g_phase_path = 1;
    for (float &x : g_phase_ms) x = 0.f;
    auto T0 = std::chrono::steady_clock::now();
    uint32_t max_id = 0;
    for (uint64_t e = 0; e < n_edges; e++) max_id = std::max(max_id, (uint32_t)keys[e]);  // b > a
    const uint32_t n = max_id + 1;
    // ref pairs (main.rs:972-980): keys are sorted, they come first
    std::vector<float> ref_w(n, 0.f);
    std::vector<uint8_t> in_ref(n, 0), bad_v(n, 0), has(n, 0);
    bool have_ref = false;
    uint64_t e0 = 0;
    for (; e0 < n_edges && (keys[e0] >> 32) == 0; e0++) {
        const uint32_t b = (uint32_t)keys[e0];
        const long long v = vals[e0];
        const long long ndif = (v + (1LL << 31)) >> 32;
        if (asref) {
            ref_w[b] = (float)(v - (ndif << 32));
            in_ref[b] = 1;
            have_ref = true;
        }
        if (ndif > 0 && !use_all_reads) bad_v[b] = 1;
    }
    // level 0: the reads (main.rs:994-1010: `dif <= -3` override, invalid reads leave, their partners stay)
    Lvl lv;
    lv.aoff_own.assign(n + 1, 0);
    for (uint64_t e = e0; e < n_edges; e++) {
        const uint32_t a = (uint32_t)(keys[e] >> 32), b = (uint32_t)keys[e];
        if (!use_all_reads && (bad_v[a] || bad_v[b])) {
            if (!bad_v[a]) has[a] = 1;
            if (!bad_v[b]) has[b] = 1;
            continue;
        }
        has[a] = has[b] = 1;
        lv.aoff_own[a + 1]++;
        lv.aoff_own[b + 1]++;
    }
    for (uint32_t v = 0; v < n; v++) lv.aoff_own[v + 1] += lv.aoff_own[v];
    lv.ato_own.resize(lv.aoff_own[n]);
    lv.aw_own.resize(lv.aoff_own[n]);
    {
        std::vector<uint32_t> cur(lv.aoff_own.begin(), lv.aoff_own.end() - 1);
        for (uint64_t e = e0; e < n_edges; e++) {  // sorted keys: every list comes out ascending
            const uint32_t a = (uint32_t)(keys[e] >> 32), b = (uint32_t)keys[e];
            if (!use_all_reads && (bad_v[a] || bad_v[b])) continue;
            const long long v = vals[e];
            const long long ndif = (v + (1LL << 31)) >> 32;  // number of disagreeing sites
            const long long sum = v - (ndif << 32);          // sum of +-1 over shared heterozygous regions
            const float w = ndif >= 3 ? -(float)ndif : (float)sum;  // main.rs:996-1002
            lv.ato_own[cur[a]] = b;
            lv.aw_own[cur[a]++] = w;
            lv.ato_own[cur[b]] = a;
            lv.aw_own[cur[b]++] = w;
        }
    }
    lv.bind_own();
    level0_vertices(lv, n, has.data());
    return phase_core(lv, n, bad_v.data(), in_ref.data(), ref_w.data(), have_ref,
                      [&]() { return phase_reads_general(keys, vals, n_edges, asref, use_all_reads); }, T0);
}

}  // namespace np2
