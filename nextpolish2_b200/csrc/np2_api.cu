// np2_api.cu — C ABI (include/np2gpu.h) and the per-contig pipeline that strings the kernels and host
// phases together.  One np2_ctx = one GPU + a compute stream (all kernels of a job), a high-priority copy stream
// (uploads, the K0 gather) and a private stream-ordered memory pool; per-run scratch comes from an arena.

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/np2gpu.h"
#include "np2_host.h"
#include "np2_inflate.cuh"
#include "np2_kernels.cuh"

using namespace np2;

namespace {
thread_local std::string g_err;

template <class F>
int guard(F f) {
    try {
        f();
        return NP2_OK;
    } catch (const np2::Error &e) {
        g_err = e.what();
        return e.code;
    } catch (const std::exception &e) {
        g_err = e.what();
        return NP2_ERR_INTERNAL;
    }
}

// Every context allocates from its OWN stream-ordered pool.  With the device's default pool shared by two contexts, a
// block freed on one stream and handed to the other makes the second stream wait for the first (the pool's
// cross-stream reuse dependencies), which serialises two contigs that are meant to be in flight together.
std::mutex g_pool_mu;
std::unordered_map<cudaStream_t, cudaMemPool_t> g_stream_pool;
cudaMemPool_t pool_of(cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    auto it = g_stream_pool.find(st);
    return it == g_stream_pool.end() ? nullptr : it->second;
}

// Per-run scratch arena.  A run makes ~150 small stream-ordered allocations (and as many frees); each is a driver call,
// and with several contigs in flight the threads queue up on the driver's lock.  Inside np2_job::run every allocation
// below kArenaMax is a pointer bump in a block that lives with the context's scratch (so in steady state a run
// allocates nothing); large buffers still come from the context's pool.
struct Arena {
    struct Block {
        uint8_t *p;
        size_t cap;
    };
    std::vector<Block> blocks;
    size_t cur = 0, off = 0, used_total = 0;
    void *bump(size_t bytes, cudaStream_t st) {
        bytes = (bytes + 255) & ~(size_t)255;
        while (cur < blocks.size() && off + bytes > blocks[cur].cap) {
            cur++;
            off = 0;
        }
        if (cur == blocks.size()) {
            Block b;
            b.cap = std::max<size_t>(bytes, 64u << 20);
            cudaMemPool_t pool = pool_of(st);
            cudaError_t e = pool ? cudaMallocFromPoolAsync((void **)&b.p, b.cap, pool, st) : cudaMallocAsync((void **)&b.p, b.cap, st);
            if (e != cudaSuccess) {
                cudaGetLastError();
                return nullptr;
            }
            blocks.push_back(b);
            off = 0;
        }
        void *r = blocks[cur].p + off;
        off += bytes;
        used_total += bytes;
        return r;
    }
    // start of a run: everything handed out before is dead (same stream, so reuse is ordered); a run that spilled
    // into several blocks gets one block of the right size next time
    void reset(cudaStream_t st) {
        if (blocks.size() > 1) {
            for (auto &b : blocks) cudaFreeAsync(b.p, st);
            blocks.clear();
            Block b;
            b.cap = used_total + used_total / 4 + (16u << 20);
            cudaMemPool_t pool = pool_of(st);
            cudaError_t e = pool ? cudaMallocFromPoolAsync((void **)&b.p, b.cap, pool, st) : cudaMallocAsync((void **)&b.p, b.cap, st);
            if (e == cudaSuccess) blocks.push_back(b);
            else cudaGetLastError();
        }
        cur = 0;
        off = 0;
        used_total = 0;
    }
    void destroy(cudaStream_t st) {
        for (auto &b : blocks) cudaFreeAsync(b.p, st);
        blocks.clear();
    }
};
// NP2_TRACE: driver calls that block for more than half a millisecond are reported with their size
// per-stage CUDA events (np2_job_get_timings): two event records per stage are two more submissions per stage; with
// several contigs in flight and the link busy every submission costs (np2_set_stage_timing(0) keeps only "total")
std::atomic<int> g_stage_events{[] {
    const char *e = getenv("NP2_STAGE_TIMING");
    return !e || atoi(e) != 0 ? 1 : 0;
}()};
thread_local std::vector<std::pair<std::string, double>> *g_trace = nullptr;
thread_local std::chrono::steady_clock::time_point g_trace_t0;
struct SlowCall {
    const char *what;
    size_t bytes;
    std::chrono::steady_clock::time_point t;
    SlowCall(const char *w, size_t b) : what(w), bytes(b), t(std::chrono::steady_clock::now()) {}
    ~SlowCall() {
        if (!g_trace) return;
        const auto now = std::chrono::steady_clock::now();
        const double ms = std::chrono::duration<double, std::milli>(now - t).count();
        if (ms < 0.5) return;
        char buf[96];
        snprintf(buf, sizeof buf, "SLOW:%s(%zuB,%.2fms)", what, bytes, ms);
        g_trace->emplace_back(buf, std::chrono::duration<double, std::milli>(now - g_trace_t0).count());
    }
};
constexpr size_t kArenaMax = 4u << 20;
thread_local Arena *g_arena = nullptr;
struct ArenaScope {
    Arena *prev;
    explicit ArenaScope(Arena *a) : prev(g_arena) { g_arena = a; }
    ~ArenaScope() { g_arena = prev; }
};

template <class T>
struct DBuf {  // stream-ordered device buffer
    T *p = nullptr;
    size_t n = 0;
    cudaStream_t s = nullptr;
    bool from_arena = false;
    DBuf() {}
    DBuf(const DBuf &) = delete;
    DBuf &operator=(const DBuf &) = delete;
    void alloc(size_t count, cudaStream_t st) {
        release();
        s = st;
        n = count;
        if (count && g_arena && count * sizeof(T) <= kArenaMax) {
            p = static_cast<T *>(g_arena->bump(count * sizeof(T), st));
            if (p) {
                from_arena = true;
                return;
            }
        }
        if (count) {
            SlowCall sc_("pool_alloc", count * sizeof(T));
            cudaMemPool_t pool = pool_of(st);
            cudaError_t e = pool ? cudaMallocFromPoolAsync((void **)&p, count * sizeof(T), pool, st)
                                 : cudaMallocAsync((void **)&p, count * sizeof(T), st);
            if (e != cudaSuccess) {
                p = nullptr;
                throw np2::Error(NP2_ERR_CUDA, std::string("cudaMallocAsync of ") + std::to_string(count) + " x " +
                                                   std::to_string(sizeof(T)) + " bytes: " + cudaGetErrorString(e));
            }
        }
    }
    void zero() {
        if (n) NP2_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s));
    }
    void upload(const T *h, size_t count) {
        SlowCall sc_("h2d", count * sizeof(T));
        NP2_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
    }
    void download(T *h, size_t count) const {
        SlowCall sc_("d2h", count * sizeof(T));
        NP2_CUDA(cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, s));
    }
    void release() {
        if (p && !from_arena) cudaFreeAsync(p, s);
        p = nullptr;
        n = 0;
        from_arena = false;
    }
    ~DBuf() { release(); }
};

template <class T>
struct PBuf {  // pinned host buffer
    T *p = nullptr;
    size_t n = 0, cap = 0;
    void resize(size_t count) {
        if (count > cap) {
            SlowCall sc_("pinned_alloc", count * sizeof(T));
            if (p) cudaFreeHost(p);
            p = nullptr;
            cap = std::max(count, cap * 2);
            NP2_CUDA(cudaMallocHost((void **)&p, cap * sizeof(T)));
        }
        n = count;
    }
    // capacity without changing n: page-locked (re)allocations stall every context of the device for milliseconds,
    // so buffers whose final size is known roughly are made large enough once
    void reserve(size_t count) {
        const size_t keep = n;
        if (count > cap) resize(count);
        n = keep;
    }
    ~PBuf() {
        if (p) cudaFreeHost(p);
    }
};

struct StageTimer {
    struct Rec {
        cudaEvent_t a, b;
        int stage;
    };
    std::vector<std::string> names;
    std::vector<float> ms;
    std::vector<uint32_t> launches;
    std::vector<Rec> recs;
    std::vector<cudaEvent_t> pool;
    cudaStream_t s = nullptr;
    int id(const char *name) {
        for (size_t i = 0; i < names.size(); i++)
            if (names[i] == name) return (int)i;
        names.push_back(name);
        ms.push_back(0);
        launches.push_back(0);
        return (int)names.size() - 1;
    }
    cudaEvent_t ev() {
        if (!pool.empty()) {
            cudaEvent_t e = pool.back();
            pool.pop_back();
            return e;
        }
        cudaEvent_t e;
        NP2_CUDA(cudaEventCreate(&e));
        return e;
    }
    // NP2_TRACE=1: wall-clock marks of every stage begin / host phase end of a run, printed to stderr when it ends
    // (shows where the HOST thread spends its time, e.g. blocked in the driver while other contexts are busy)
    std::vector<std::pair<std::string, double>> trace;
    std::chrono::steady_clock::time_point trace_t0;
    static bool tracing() {
        static const bool on = [] {
            const char *e = getenv("NP2_TRACE");
            return e && atoi(e) != 0;
        }();
        return on;
    }
    void mark(const char *what) {
        if (tracing())
            trace.emplace_back(what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - trace_t0).count());
    }
    void trace_begin() {
        trace.clear();
        trace_t0 = std::chrono::steady_clock::now();
        g_trace = tracing() ? &trace : nullptr;
        g_trace_t0 = trace_t0;
    }
    void trace_dump(const void *tag) {
        if (!tracing() || trace.empty()) return;
        std::string out = "[np2 trace " + std::to_string((uintptr_t)tag % 100000) + "]";
        double prev = 0;
        char buf[160];
        for (auto &t : trace) {
            if (t.second - prev >= 0.05 || t.first.compare(0, 5, "SLOW:") == 0) {
                snprintf(buf, sizeof buf, " %s@%.2f(+%.2f)", t.first.c_str(), t.second, t.second - prev);
                out += buf;
            }
            prev = t.second;
        }
        fprintf(stderr, "%s\n", out.c_str());
    }
    int begin(const char *name, uint32_t n_launch) {
        mark(name);
        if (!g_stage_events.load(std::memory_order_relaxed) && strcmp(name, "total") != 0) {
            launches[id(name)] += n_launch;
            return -1;
        }
        Rec r;
        r.stage = id(name);
        r.a = ev();
        r.b = ev();
        launches[r.stage] += n_launch;
        NP2_CUDA(cudaEventRecord(r.a, s));
        recs.push_back(r);
        return (int)recs.size() - 1;
    }
    void end(int h) {
        if (h >= 0) NP2_CUDA(cudaEventRecord(recs[h].b, s));
    }
    void collect() {  // after a stream sync
        for (auto &r : recs) {
            float t = 0;
            cudaEventElapsedTime(&t, r.a, r.b);
            ms[r.stage] += t;
            pool.push_back(r.a);
            pool.push_back(r.b);
        }
        recs.clear();
    }
    // host phases: wall clock, reported as "host:<name>"
    std::chrono::steady_clock::time_point h0;
    void hbegin() { h0 = std::chrono::steady_clock::now(); }
    void hend(const char *name) {
        mark(name);
        auto t = std::chrono::steady_clock::now();
        ms[id(name)] += std::chrono::duration<float, std::milli>(t - h0).count();
        h0 = t;
    }
    void reset(bool upload_too = false) {  // "upload:*" stages belong to np2_job_create/upload and survive a run
        for (size_t i = 0; i < names.size(); i++)
            if (upload_too || names[i].compare(0, 7, "upload:") != 0) {
                ms[i] = 0.f;
                launches[i] = 0u;
            }
    }
    ~StageTimer() {
        for (auto e : pool) cudaEventDestroy(e);
    }
};
}  // namespace

// Host-side buffers a job needs, pooled per context: pinned allocations (cudaMallocHost) and first-touch page
// faults of fresh vectors cost milliseconds each, far more than the kernels they feed.
constexpr int kHintSlots = 4;
// sizes of the last pass of one kind (slot = iteration the pass started at): the capacities of the next one
struct Hints {
    bool valid = false;
    bool in_pass = false;  // the pass under way has already written its counts here (later fetches of the pass merge)
    uint32_t L = 0;
    uint64_t cols = 0;
    CountsHost h = CountsHost();
};
struct Respeculate {};  // a speculative pass met a count above its capacity: repeat it in exact mode
struct JobScratch {
    Ingest ing;
    Hints hints[kHintSlots];
    ScanPool scan_pool;
    uint32_t *d_counts = nullptr;  // CountsDev block (256 bytes)
    PBuf<uint8_t> p_counts;        // its pinned mirror (+ two words for the FASTA header positions)
    PBuf<uint8_t> p_h2d;           // staging area of the small uploads of a run (Uploader)
    size_t h2d_used = 0;
    std::vector<uint8_t> tseq, h_seeds, h_rech_pool, h_win;
    Patched patch;
    PBuf<uint8_t> res_base, p_cbase, p_cflags, p_stage, p_stage2, p_seq_stage;
    PBuf<uint32_t> p_cpos;
    std::vector<uint64_t> cseq_off, cseq_start, cspan_off;  // span / SEQ offsets in the compact device blob, span sources
    std::vector<uint32_t> cspan_bytes;
    cudaEvent_t seq_ev[2] = {nullptr, nullptr};
    cudaEvent_t ev_alloc = nullptr, ev_copied = nullptr, ev_t0 = nullptr, ev_t1 = nullptr, ev_k0 = nullptr, ev_trim = nullptr;
    PBuf<uint8_t> p_trim;
    Arena arena;  // per-run device scratch (small allocations)
    PBuf<uint8_t> p_up_stage, p_seq_args, p_phase;
    StageTimer timer;
    ~JobScratch() {
        for (auto e : seq_ev)
            if (e) cudaEventDestroy(e);
        if (ev_alloc) cudaEventDestroy(ev_alloc);
        if (ev_copied) cudaEventDestroy(ev_copied);
        if (ev_t0) cudaEventDestroy(ev_t0);
        if (ev_t1) cudaEventDestroy(ev_t1);
        if (ev_k0) cudaEventDestroy(ev_k0);
        if (ev_trim) cudaEventDestroy(ev_trim);
    }
};
// bump allocator over a pinned staging buffer: many small device arrays come back with one synchronisation and
// without the implicit host-side staging of pageable destinations
struct Stager {
    PBuf<uint8_t> &buf;
    cudaStream_t s;
    size_t used = 0;
    Stager(PBuf<uint8_t> &b, cudaStream_t st, size_t cap) : buf(b), s(st) { buf.resize(std::max<size_t>(cap, 64)); }
    template <class T>
    T *fetch(const T *dev, size_t n) {
        used = (used + 15) & ~(size_t)15;
        if (used + n * sizeof(T) > buf.cap) throw np2::Error(NP2_ERR_INTERNAL, "staging buffer overflow");
        T *dst = reinterpret_cast<T *>(buf.p + used);
        if (n) NP2_CUDA(cudaMemcpyAsync(dst, dev, n * sizeof(T), cudaMemcpyDeviceToHost, s));
        used += n * sizeof(T);
        return dst;
    }
};
// Host -> device copies of small host arrays through a page-locked staging area: a cudaMemcpyAsync from pageable memory
// blocks the calling thread inside the driver (for tens of milliseconds when other contexts are busy), a copy from
// page-locked memory is just enqueued.  The area is handed out by a bump pointer and recycled when a run starts.
struct Uploader {
    PBuf<uint8_t> &buf;
    size_t &used;
    cudaStream_t s;
    template <class T>
    void put(T *dev, const T *host, size_t count) {
        const size_t bytes = count * sizeof(T);
        if (!bytes) return;
        const size_t at = (used + 15) & ~(size_t)15;
        if (at + bytes > buf.cap) {  // does not fit: plain (blocking) copy
            NP2_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, s));
            return;
        }
        memcpy(buf.p + at, host, bytes);
        NP2_CUDA(cudaMemcpyAsync(dev, buf.p + at, bytes, cudaMemcpyHostToDevice, s));
        used = at + bytes;
    }
};
struct np2_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // H2D of the per-read arrays, concurrent with the K0 gather
    cudaMemPool_t pool = nullptr;
    uint64_t pool_warm = 0;  // bytes the pool has been grown to in one step (np2_job_create)
    int refs = 1;  // tables and jobs keep their context alive (np2_ctx_destroy only drops the caller's reference)
    std::vector<JobScratch *> scratch_pool;
    PBuf<uint8_t> p_infl_in, p_infl_args;  // np2_bgzf_inflate: staged compressed bytes, member table
    JobScratch *take_scratch() {
        if (scratch_pool.empty()) return new JobScratch();
        JobScratch *sc = scratch_pool.back();
        scratch_pool.pop_back();
        return sc;
    }
};
static void ctx_release(np2_ctx *ctx) {
    if (--ctx->refs > 0) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (JobScratch *sc : ctx->scratch_pool) {
        sc->arena.destroy(ctx->stream);
        sc->scan_pool.destroy(ctx->stream);
        if (sc->d_counts) cudaFree(sc->d_counts);
        delete sc;
    }
    cudaStreamSynchronize(ctx->stream);
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        g_stream_pool.erase(ctx->stream);
    }
    cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
    delete ctx;
}

struct np2_table {
    np2_ctx *ctx = nullptr;
    TableDev dev;
    uint64_t bytes = 0;
};
static void ctx_release(np2_ctx *ctx);

static void table_alloc(np2_ctx *ctx, np2_table *t, uint32_t k, uint64_t n, uint64_t max_sub) {
    t->ctx = ctx;
    ctx->refs++;
    t->dev.k = k;
    t->dev.n = n;
    // buckets of 4 slots, load factor <= 0.6 in the fullest sub-table
    uint64_t nb = (uint64_t)((double)max_sub / (kBucketSlots * 0.6)) + 2;
    if (nb >= (1ull << 32)) throw np2::Error(NP2_ERR_UNSUPPORTED, "table too large");
    t->dev.nb = (uint32_t)nb;
    t->bytes = 1024ull * nb * kBucketSlots * 8;
    NP2_CUDA(cudaMalloc((void **)&t->dev.slots, t->bytes));
    NP2_CUDA(cudaMemsetAsync(t->dev.slots, 0, t->bytes, ctx->stream));
}

struct np2_job {
    np2_ctx *ctx = nullptr;
    JobScratch *sc = nullptr;
    np2_opts opt;
    std::vector<np2_table *> tables;  // sorted by k
    std::vector<uint8_t> &tseq;  // host copy only for contigs below min_ctg_len (echoed back unchanged)
    uint32_t L = 0;
    const uint8_t *bam = nullptr;
    uint64_t bam_len = 0;
    DBuf<uint8_t> d_records;         // np2_job_create_bgzf: the inflated members (bam points into them)
    bool records_on_device = false;
    Ingest &ing;
    bool uploaded = false;
    uint64_t seq_blob_bytes = 0;
    int seq_path = 0;  // 1 = gathered by the device from page-locked records, 2 = compacted by host threads,
                       // 3 = gathered device to device from records inflated there (np2_job_create_bgzf)
    explicit np2_job(np2_ctx *c)
        : ctx(c), sc(c->take_scratch()), tseq(sc->tseq), ing(sc->ing), res_base(sc->res_base), p_cpos(sc->p_cpos),
          p_cbase(sc->p_cbase), p_cflags(sc->p_cflags), timer(sc->timer), h_seeds(sc->h_seeds), h_win(sc->h_win), res_patch(sc->patch),
          h_rech_pool(sc->h_rech_pool) {}
    ~np2_job() { ctx->scratch_pool.push_back(sc); }

    // device inputs
    DBuf<uint8_t> d_ref, d_code, d_blob, d_nib, d_blank;
    DBuf<int> d_bad;  // the contig holds a byte the reference cannot index
    DBuf<uint32_t> d_refpk;
    DBuf<uint32_t> d_pos, d_op_off, d_ncols, d_ck_off, d_ncig;
    DBuf<Op> d_ops;
    DBuf<uint64_t> d_seq_off, d_nib_off;
    DBuf<uint32_t> d_ts, d_te, d_n, d_shift, d_ck_tpos, d_ck_read;
    DBuf<uint16_t> d_ck_delta, d_blk_op;
    ReadsDev R;

    // host state
    std::vector<uint32_t> h_ts, h_te, h_n;
    std::vector<uint8_t> h_blank;          // per candidate read
    std::vector<int32_t> as_read;          // alignseq index -> candidate read (-1 = ref)
    uint32_t rec_cap_hint = 0;             // 3-mer records of the last pileup (sizes the next one-pass emit)
    std::vector<uint32_t> h_as_pos, h_as_te;  // record pos / last column of every alignseq (pair-accumulator windows)
    DBuf<uint64_t> d_pair_off;
    DBuf<uint32_t> d_as_pos;               // record position of every alignseq (device copy of h_as_pos)
    DBuf<uint32_t> d_first_ge;             // read window of every pileup stripe (np2_kernels.cu stripe_reads)
    DBuf<uint32_t> d_blk_odd;              // one bit per 32-column block: not all reference (np2_kernels.cu block_flags)
    DBuf<uint32_t> d_odd_off, d_odd_list;  // per pileup stripe: the flagged blocks that can touch it (np2_kernels.cu k_stripe_odd)
    std::vector<uint32_t> read_order;      // candidate read -> alignseq index (0 = not kept)

    // result: bases always; positions are materialised on request (np2_job_get_consensus with pos != NULL)
    PBuf<uint8_t> &res_base;
    std::vector<uint32_t> res_pos;
    DBuf<uint32_t> jd_cpos;                  // DP consensus of the last iteration stays on the device
    DBuf<uint32_t> jd_reg_start, jd_reg_a, jd_reg_b, jd_reg_len;  // final regions (r order) for lazy positions
    uint32_t res_nreg = 0;
    bool res_sparse = false;                 // res_patch only holds the regions the re-check looked at
    DBuf<uint8_t> jd_cbase, jd_cflags;
    uint32_t res_N = 0;
    std::vector<uint8_t> &h_seeds, &h_rech_pool, &h_win;
    uint32_t res_first = 0, res_last = 0;
    bool res_pos_valid = false;
    Patched &res_patch;                      // kept so that positions can be produced lazily (pooled with the scratch)
    PBuf<uint32_t> &p_cpos;
    PBuf<uint8_t> &p_cbase, &p_cflags;
    DBuf<uint32_t> d_order;
    uint32_t max_span = 0;
    StageTimer &timer;
    uint64_t h2d = 0, d2h = 0, n_launch = 0, n_probes = 0;
    uint64_t stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // np2_job_get_stats
    uint64_t n_spec_ok = 0, n_respec = 0, n_sync = 0;  // speculative passes, passes repeated in exact mode, host syncs of the run
    bool spec = false;        // the pass under way sizes everything from capacities and does not read counts back
    CountsHost caps = CountsHost();
    CountsHost *hc = nullptr;  // pinned mirror of the device counts
    CountsDev cd;
    uint64_t pair_slots = 0;  // slots of the dense pair accumulator (depends on the reads only)
    std::string timing_names;

    // dumps
    int32_t dump_iter = -1;
    std::vector<int32_t> d_rec_idx;
    std::vector<uint32_t> dm_ts, dm_te;
    std::vector<uint64_t> dm_nib_off;
    std::vector<uint8_t> dm_nib, dm_blank;
    std::vector<uint64_t> dm_msa_off;
    std::vector<uint16_t> dm_msa_bases, dm_msa_delta;
    std::vector<uint32_t> dm_msa_count, dm_msa_besti;
    std::vector<uint32_t> dm_dp_pos;
    std::vector<uint8_t> dm_dp_base, dm_dp_flags;
    std::vector<uint32_t> dm_reg_start, dm_reg_end;
    std::vector<uint8_t> dm_reg_lable;
    std::vector<uint64_t> dm_can_roff, dm_can_kmer, dm_can_seq_off;
    std::vector<uint32_t> dm_can_order;
    std::vector<uint16_t> dm_can_kscore;
    std::vector<uint8_t> dm_can_seq;
    std::vector<uint32_t> dm_dropped;
    std::vector<uint64_t> dm_pair_key;
    std::vector<int64_t> dm_pair_val;

    void send_contig(const uint8_t *tseq_host);
    void send_seq();
    void enqueue_arrays();
    void upload();
    void run(int32_t dump_iter);
    void ingest_finish();
    uint32_t iteration(uint32_t iter0);
    uint32_t iteration_pass(uint32_t iter0, Hints &hint);
    void fetch_counts();
    uint32_t cnt_get(int idx);
    void segment_end();
};

/* ================================================================= pipeline */

// One K0 at a time per device.  A K0 keeps ~0.5 MB of reads outstanding on the link; every other transfer — the
// command fetches of every kernel launch of the contigs that are being polished meanwhile, their count read-backs —
// queues behind those reads, and with two or three K0s of different contexts running at once a kernel launch goes
// from ~8 to ~40 us (profiles/r02w_diag_probe.json).  The link is the shared resource anyway, so the K0s of a device
// are chained through events: each waits for the one enqueued before it, on the device, without blocking a host thread.
struct K0Chain {
    std::mutex mu;
    static constexpr int kRing = 64;
    cudaEvent_t ev[kRing] = {};
    uint64_t n = 0;
};
K0Chain g_k0_chain[16];
static const bool g_k0_serial = [] {
    const char *e = getenv("NP2_K0_SERIAL");
    return !e || atoi(e) != 0;
}();
void k0_chain_enter(int device, cudaStream_t c2) {  // before the K0 launch; returns with the chain's mutex HELD
    K0Chain &ch = g_k0_chain[device & 15];
    ch.mu.lock();
    if (ch.n) NP2_CUDA(cudaStreamWaitEvent(c2, ch.ev[(ch.n - 1) % K0Chain::kRing], 0));
}
void k0_chain_leave(int device, cudaStream_t c2) {  // after the K0 launch
    K0Chain &ch = g_k0_chain[device & 15];
    cudaEvent_t &e = ch.ev[ch.n % K0Chain::kRing];
    cudaError_t err = cudaSuccess;
    if (!e) err = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    if (err == cudaSuccess) err = cudaEventRecord(e, c2);
    ch.n++;
    ch.mu.unlock();
    if (err != cudaSuccess) throw np2::Error(NP2_ERR_CUDA, cudaGetErrorString(err));
}

// The contig goes up straight from the caller's buffer (asynchronously when it is page-locked).
void np2_job::send_contig(const uint8_t *tseq_host) {
    cudaStream_t s = ctx->stream;
    timer.s = s;
    timer.reset(true);
    int h = timer.begin("upload:contig", 0);
    d_ref.alloc(L, s);
    d_ref.upload(tseq_host, L);
    timer.end(h);
    d_code.alloc(L, s);
    d_refpk.alloc(L / 8 + 8, s);
    d_bad.alloc(1, s);
    d_bad.zero();
    h = timer.begin("upload:ref_codes", 2);  // SEQ_NUM codes of the contig: input preparation, once per job
    ref_codes(d_ref.p, L, d_code.p, d_refpk.p, d_bad.p, s);
    timer.end(h);
    h2d += L;
}

// Only the 4-bit SEQ fields of the kept records go to the device (a third of the record bytes for HiFi BAMs with
// QUAL): compact blob, one 16-byte aligned slot per read that keeps the source's misalignment.
//  * page-locked caller buffer (cudaHostAlloc / cudaHostRegister): K0 gathers over PCIe from the mapped records
//  * pageable buffer: host threads compact into a pooled pinned ring, one DMA per round
void np2_job::send_seq() {
    cudaStream_t s = ctx->stream;
    const uint32_t n = (uint32_t)ing.pos.size();
    const uint8_t *src = bam;
    seq_path = 2;
    if (bam_len) {
        cudaPointerAttributes at;
        cudaError_t e = cudaPointerGetAttributes(&at, bam);
        if (e != cudaSuccess) cudaGetLastError();  // older runtimes flag plain malloc memory as an error
        else if (at.type == cudaMemoryTypeHost && at.devicePointer) {
            src = static_cast<const uint8_t *>(at.devicePointer);
            seq_path = 1;
        } else if (at.type == cudaMemoryTypeDevice) {  // records inflated on this device (np2_job_create_bgzf)
            seq_path = 3;
        }
    }
    // One span per read: its raw CIGAR words and, right behind them in the record, its SEQ bytes.  co[] = where the
    // span lands in the blob, sq[] = where its SEQ part starts (what the kernels index), span_off / span_bytes = the source.
    std::vector<uint64_t> &co = sc->cseq_off, &sq = sc->cseq_start, &span_off = sc->cspan_off;
    std::vector<uint32_t> &span_bytes = sc->cspan_bytes;
    co.resize(n);
    sq.resize(n);
    span_off.resize(n);
    span_bytes.resize(n);
    uint64_t D = 0;
    for (uint32_t r = 0; r < n; r++) {
        const uint64_t cig = ing.host_ops ? 0 : 4ull * ing.n_cig[r];
        span_off[r] = ing.seq_off[r] - cig;
        span_bytes[r] = (uint32_t)(ing.seq_bytes[r] + cig);
        const uint64_t mis = (uintptr_t)(src + span_off[r]) & 15;
        co[r] = D + mis;
        sq[r] = co[r] + cig;
        D += ((mis + span_bytes[r] + 15) & ~15ull) + 32;  // slack: the pack kernel reads whole words past the end
    }
    seq_blob_bytes = D;
    d_blob.alloc(D + 64, s);
    d_seq_off.alloc(std::max<size_t>(n, 1), s);
    // offsets go through a pinned staging buffer: a pageable cudaMemcpyAsync would block this thread
    sc->p_seq_args.resize((size_t)n * 28 + 64);
    uint8_t *st = sc->p_seq_args.p;
    if (n) {
        np2::copy_streaming(st, sq.data(), (size_t)n * 8);
        np2::store_fence();
        NP2_CUDA(cudaMemcpyAsync(d_seq_off.p, st, (size_t)n * 8, cudaMemcpyHostToDevice, s));
    }
    h2d += (uint64_t)n * 8;
    if (!n) return;
    if (seq_path == 1 || seq_path == 3) {
        DBuf<uint64_t> d_src_off, d_dst_off;
        DBuf<uint32_t> d_nbytes;
        d_src_off.alloc(n, s);
        d_dst_off.alloc(n, s);
        d_nbytes.alloc(n, s);
        np2::copy_streaming(st + (size_t)n * 8, span_off.data(), (size_t)n * 8);
        np2::copy_streaming(st + (size_t)n * 16, co.data(), (size_t)n * 8);
        np2::copy_streaming(st + (size_t)n * 24, span_bytes.data(), (size_t)n * 4);
        np2::store_fence();
        NP2_CUDA(cudaMemcpyAsync(d_src_off.p, st + (size_t)n * 8, (size_t)n * 8, cudaMemcpyHostToDevice, s));
        NP2_CUDA(cudaMemcpyAsync(d_dst_off.p, st + (size_t)n * 16, (size_t)n * 8, cudaMemcpyHostToDevice, s));
        NP2_CUDA(cudaMemcpyAsync(d_nbytes.p, st + (size_t)n * 24, (size_t)n * 4, cudaMemcpyHostToDevice, s));
        // K0 runs on the context's high-priority copy stream: its CTAs only wait on PCIe, but they must be RESIDENT
        // to keep enough loads in flight; when another contig's kernels fill the SMs, a normal-priority K0 gets its
        // CTAs scheduled late and the link idles.  The main stream joins again before the offsets are freed.
        cudaStream_t c2 = ctx->copy_stream;
        if (!sc->ev_k0) NP2_CUDA(cudaEventCreateWithFlags(&sc->ev_k0, cudaEventDisableTiming));
        NP2_CUDA(cudaEventRecord(sc->ev_k0, s));  // allocations + offset uploads above
        NP2_CUDA(cudaStreamWaitEvent(c2, sc->ev_k0, 0));
        timer.s = c2;
        int h = timer.begin("upload:seq_gather", 1);
        if (g_k0_serial && seq_path == 1) {  // one K0 at a time on the LINK; a device-to-device gather does not use it
            k0_chain_enter(ctx->device, c2);
            try {
                gather_seq(src, d_src_off.p, d_dst_off.p, d_nbytes.p, d_blob.p, n, c2, seq_path == 3);
            } catch (...) {
                g_k0_chain[ctx->device & 15].mu.unlock();
                throw;
            }
            k0_chain_leave(ctx->device, c2);
        } else {
            gather_seq(src, d_src_off.p, d_dst_off.p, d_nbytes.p, d_blob.p, n, c2, seq_path == 3);
        }
        timer.end(h);
        timer.s = s;
        NP2_CUDA(cudaEventRecord(sc->ev_k0, c2));
        NP2_CUDA(cudaStreamWaitEvent(s, sc->ev_k0, 0));
        h2d += (uint64_t)n * 20;
        if (seq_path == 1)
            for (uint32_t r = 0; r < n; r++) h2d += span_bytes[r];
        return;  // scratch is freed in stream order, after the kernel
    }
    // pageable source: rounds of <= kRound bytes through two halves of a pinned ring
    const uint64_t kRound = 64ull << 20;
    uint64_t max_slot = 0;
    for (uint32_t r = 0; r < n; r++) max_slot = std::max<uint64_t>(max_slot, (uint64_t)span_bytes[r] + 64);
    const uint64_t half = std::max(kRound, max_slot);
    sc->p_seq_stage.resize(2 * half);
    for (auto &ev : sc->seq_ev)
        if (!ev) NP2_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    const unsigned T = np2::host_threads();
    uint32_t r0 = 0;
    for (uint32_t round = 0; r0 < n; round++) {
        const uint64_t D0 = co[r0] & ~15ull;
        uint32_t r1 = r0 + 1;
        auto slot_end = [&](uint32_t r) { return (r + 1 < n) ? (co[r + 1] & ~15ull) : D; };
        while (r1 < n && slot_end(r1) - D0 <= half) r1++;
        const uint64_t D1 = slot_end(r1 - 1);
        uint8_t *buf = sc->p_seq_stage.p + (round & 1) * half;
        if (round >= 2) NP2_CUDA(cudaEventSynchronize(sc->seq_ev[round & 1]));
        auto work = [&](unsigned ti) {
            const uint32_t b = r0 + (uint64_t)(r1 - r0) * ti / T, e = r0 + (uint64_t)(r1 - r0) * (ti + 1) / T;
            for (uint32_t r = b; r < e; r++) np2::copy_streaming(buf + (co[r] - D0), bam + span_off[r], span_bytes[r]);
            np2::store_fence();  // the DMA reads this ring next: keep it out of the cores' caches
        };
        if (r1 - r0 < 256) {
            for (unsigned ti = 0; ti < T; ti++) work(ti);
        } else {
            np2::parallel_for(T, work);
        }
        NP2_CUDA(cudaMemcpyAsync(d_blob.p + D0, buf, D1 - D0, cudaMemcpyHostToDevice, s));
        NP2_CUDA(cudaEventRecord(sc->seq_ev[round & 1], s));
        h2d += D1 - D0;
        r0 = r1;
    }
}

// Everything but the SEQ bytes: per-read scalars and the CIGAR op arrays.  Enqueued by np2_job_create right after
// the parse, on the context's COPY stream, so that these DMA transfers run while the main stream's K0 kernel pulls
// the SEQ bytes over the same link and the host never blocks on a pageable copy (the small arrays are staged through
// one pinned buffer).  Device buffers are allocated on the main stream; the two streams meet through two events.
void np2_job::enqueue_arrays() {
    cudaStream_t s = ctx->stream, c2 = ctx->copy_stream;
    const uint32_t n = (uint32_t)ing.pos.size();
    timer.hbegin();
    const size_t no = std::max<size_t>(ing.n_ops, 1);
    d_pos.alloc(std::max(n, 1u), s);
    d_op_off.alloc(n + 1, s);
    d_ncols.alloc(std::max(n, 1u), s);
    d_ck_off.alloc(n + 1, s);
    d_nib_off.alloc(n + 1, s);
    d_ops.alloc(no, s);
    const uint32_t nck = ing.ck_off.back();
    d_nib.alloc(ing.nib_off.back() + 16, s);
    d_ts.alloc(std::max(n, 1u), s);
    d_te.alloc(std::max(n, 1u), s);
    d_n.alloc(std::max(n, 1u), s);
    d_shift.alloc(std::max(n, 1u), s);
    d_ck_tpos.alloc(std::max(nck, 1u), s);
    d_ck_delta.alloc(std::max(nck, 1u), s);
    d_ck_read.alloc(std::max(nck, 1u), s);
    d_blk_op.alloc(std::max(nck, 1u), s);
    d_blank.alloc(std::max(n, 1u), s);
    d_ncig.alloc(std::max(n, 1u), s);
    if (!sc->ev_alloc) {
        NP2_CUDA(cudaEventCreateWithFlags(&sc->ev_alloc, cudaEventDisableTiming));
        NP2_CUDA(cudaEventCreateWithFlags(&sc->ev_copied, cudaEventDisableTiming));
    }
    NP2_CUDA(cudaEventRecord(sc->ev_alloc, s));
    NP2_CUDA(cudaStreamWaitEvent(c2, sc->ev_alloc, 0));
    // small arrays -> one pinned staging buffer -> device
    const size_t small = (size_t)n * 12 + (size_t)(n + 1) * 20 + 256;
    sc->p_up_stage.resize(small);
    uint8_t *st = sc->p_up_stage.p;
    auto stage = [&](void *dev, const void *host, size_t bytes) {
        if (!bytes) return;
        np2::copy_streaming(st, host, bytes);
        NP2_CUDA(cudaMemcpyAsync(dev, st, bytes, cudaMemcpyHostToDevice, c2));
        st += (bytes + 15) & ~(size_t)15;
        h2d += bytes;
    };
    stage(d_pos.p, ing.pos.data(), (size_t)n * 4);
    stage(d_ncols.p, ing.ncols.data(), (size_t)n * 4);
    stage(d_op_off.p, ing.op_off.data(), (size_t)(n + 1) * 4);
    stage(d_ck_off.p, ing.ck_off.data(), (size_t)(n + 1) * 4);
    stage(d_nib_off.p, ing.nib_off.data(), (size_t)(n + 1) * 8);
    stage(d_ncig.p, ing.n_cig.data(), (size_t)n * 4);
    if (!sc->ev_t0) {
        NP2_CUDA(cudaEventCreate(&sc->ev_t0));
        NP2_CUDA(cudaEventCreate(&sc->ev_t1));
    }
    np2::store_fence();
    NP2_CUDA(cudaEventRecord(sc->ev_t0, c2));
    {  // the op records go up chunk by chunk from the page-locked segment arrays they were parsed into
        size_t w = 0;
        for (const Ingest::OpChunk &c : ing.op_chunks) {
            NP2_CUDA(cudaMemcpyAsync(d_ops.p + w, c.ops, c.n * sizeof(Op), cudaMemcpyHostToDevice, c2));
            w += c.n;
        }
        if (!ing.op_chunks.empty()) h2d += (uint64_t)ing.n_ops * 16;
    }
    NP2_CUDA(cudaEventRecord(sc->ev_t1, c2));
    NP2_CUDA(cudaEventRecord(sc->ev_copied, c2));
    R.n_reads = n;
    R.pos = d_pos.p;
    R.op_off = d_op_off.p;
    R.ncols = d_ncols.p;
    R.nib_off = d_nib_off.p;
    R.ck_off = d_ck_off.p;
    R.ops = reinterpret_cast<const uint4 *>(d_ops.p);
    R.t_s = d_ts.p;
    R.t_e = d_te.p;
    R.n = d_n.p;
    R.shift = d_shift.p;
    R.nib = d_nib.p;
    R.ck_tpos = d_ck_tpos.p;
    R.ck_delta = d_ck_delta.p;
    R.ck_read = d_ck_read.p;
    R.blk_op = d_blk_op.p;
    timer.hend("upload:host_enqueue_arrays");
}

void np2_job::upload() {
    cudaStream_t s = ctx->stream;
    R.seq_off = d_seq_off.p;
    R.blob = d_blob.p;
    timer.hbegin();
    NP2_CUDA(cudaStreamSynchronize(s));  // the main stream has waited for the copy stream (np2_job_create)
    timer.hend("upload:host_wait");
    d_records.release();  // np2_job_create_bgzf: the spans have been gathered out of the inflated records
    timer.collect();
    if (sc->ev_t0 && L >= opt.min_ctg_len) {
        float t = 0;
        if (cudaEventElapsedTime(&t, sc->ev_t0, sc->ev_t1) == cudaSuccess) timer.ms[timer.id("upload:ops_on_copy_stream")] += t;
    }
    uploaded = true;
}

// after K1: which candidate reads become alignseqs (main.rs:1800-1813), clip filter (main.rs:531-574)
void np2_job::ingest_finish() {
    const uint32_t n = R.n_reads;
    h_blank.assign(std::max(n, 1u), 1);
    read_order.assign(n, 0);
    as_read.clear();
    as_read.push_back(-1);
    std::vector<uint32_t> as_ts{0}, as_te{L - 1};
    std::vector<uint8_t> as_lab{0};
    // "Unsorted input file!" (main.rs:1753-1756): every record is compared with the last PUSHED read
    {
        int64_t pre_tid = 0, pre_pos = 0;
        size_t c = 0;
        for (size_t rec = 0; rec < ing.all_tid.size(); rec++) {
            if (!(ing.all_tid[rec] > pre_tid || (int64_t)ing.all_pos[rec] >= pre_pos))
                throw np2::Error(NP2_ERR_FORMAT, "Unsorted input file!");
            if (c < n && ing.rec_idx[c] == (int32_t)rec) {
                if (h_n[c] > opt.min_map_len && !(ing.is_clip[c] && L < 500000)) {
                    pre_tid = ing.all_tid[rec];
                    pre_pos = ing.all_pos[rec];
                }
                c++;
            }
        }
    }
    for (uint32_t i = 0; i < n; i++) {
        if (h_n[i] <= opt.min_map_len) continue;
        if (ing.is_clip[i] && L < 500000) continue;
        read_order[i] = (uint32_t)as_read.size();
        as_read.push_back((int32_t)i);
        as_ts.push_back(h_ts[i]);
        as_te.push_back(h_te[i]);
        as_lab.push_back(ing.is_clip[i]);
        h_blank[i] = 0;
    }
    // inputs of the pair-accumulator windows (np2_geno.cu k_pair_windows): alignseqs are in position order, so a later
    // read y can only share a region with x when it starts before x ends
    h_as_pos.assign(as_read.size(), 0);
    for (size_t a = 1; a < as_read.size(); a++) h_as_pos[a] = ing.pos[as_read[a]];
    h_as_te = as_te;
    // merged [t_s + 50, t_e - 50] of unlabelled reads; labelled reads inside a range are blanked
    std::vector<std::pair<uint32_t, uint32_t>> ranges;
    uint32_t s = 0, e = 0;
    for (size_t a = 0; a < as_read.size(); a++) {
        if (as_lab[a]) continue;
        const uint32_t x = as_ts[a] + 50, y = as_te[a] - 50;
        if (s == e) {
            s = x;
            e = y;
        } else if (x > e) {
            ranges.emplace_back(s, e);
            s = x;
            e = y;
        } else if (e < y) {
            e = y;
        }
    }
    if (s != e) ranges.emplace_back(s, e);
    for (size_t a = 0; a < as_read.size(); a++) {
        if (!as_lab[a]) continue;
        for (auto &r : ranges) {
            if (r.first <= as_ts[a] && as_te[a] <= r.second) {
                h_blank[as_read[a]] = 1;
                break;
            } else if (as_te[a] < r.first) {
                break;
            }
        }
    }
}

// One pass = the device work of iteration iter0 of the reference's loop (main.rs:1819-1836) plus the host phase that
// follows it.  When a non-final iteration blanks no read, the next iteration would rebuild exactly the same Msa,
// consensus, regions and candidates, so it is served from the state already on the device instead of being
// recomputed.  Returns the next iteration that needs a fresh build.
//
// Sizes live on the device (CountsDev).  In SPECULATIVE mode the capacities of every buffer and grid come from the last
// pass of the same kind (scaled to this contig), nothing is read back until the host needs data, and a count that
// exceeds its capacity aborts the rest of the pass on the device: the pass is then repeated in EXACT mode, where every
// count is read back before the buffers that depend on it are sized (one synchronisation per count, as a first run
// has to do anyway).
uint32_t np2_job::iteration(uint32_t iter0) {
    const uint64_t probes0 = n_probes, built0 = stats[7];
    const size_t dropped0 = dm_dropped.size();
    Hints &hint = sc->hints[std::min<uint32_t>(iter0, kHintSlots - 1)];
    static const bool enabled = [] {
        const char *e = getenv("NP2_SPECULATE");
        return !e || atoi(e) != 0;
    }();
    hint.in_pass = false;
    if (enabled && dump_iter < 0 && hint.valid) {
        // scale the remembered counts to this contig; 25 % + 1024 of slack
        const double sc_l = hint.L ? (double)L / hint.L : 1.0, sc_c = hint.cols ? (double)ing.total_cols / hint.cols : 1.0;
        const double scale = std::max(sc_l, sc_c) * 1.25;
        for (int i = 0; i < C_COUNT; i++) caps.c[i] = (uint32_t)std::min<double>(hint.h.c[i] * scale + 1024.0, 4.0e9);
        for (int i = 0; i < Q_COUNT; i++) caps.q[i] = (unsigned long long)(hint.h.q[i] * scale) + 65536;
        try {
            spec = true;
            const uint32_t r = iteration_pass(iter0, hint);
            n_spec_ok++;
            return r;
        } catch (const Respeculate &) {
            n_probes = probes0;
            stats[7] = built0;
            dm_dropped.resize(dropped0);
            n_respec++;
        }
    }
    spec = false;
    return iteration_pass(iter0, hint);
}

void np2_job::fetch_counts() {
    cudaStream_t s = ctx->stream;
    NP2_CUDA(cudaMemcpyAsync(hc, cd.c, sizeof(CountsHost), cudaMemcpyDeviceToHost, s));
    NP2_CUDA(cudaStreamSynchronize(s));
    n_sync++;
}
// capacity of a count in speculative mode, its exact value (one synchronisation) in exact mode
uint32_t np2_job::cnt_get(int idx) {
    if (spec) return caps.c[idx];
    fetch_counts();
    return hc->c[idx];
}
// end of the speculative stretch: everything enqueued so far has its counts in *hc; from here on the pass is exact
void np2_job::segment_end() {
    fetch_counts();
    if (spec && hc->c[C_ABORT]) throw Respeculate();
    spec = false;
}

uint32_t np2_job::iteration_pass(uint32_t iter0, Hints &hint) {
    cudaStream_t s = ctx->stream;
    const uint32_t n_reads = R.n_reads;
    const uint32_t n_blocks = ing.ck_off.back();
    ScanPool &sp = sc->scan_pool;
    int h;
    sp.begin(s);
    counts_init(cd, s);

    /* ---------------- K2: pileup */
    DBuf<int32_t> d_cover;
    d_cover.alloc(L + 1, s);
    DBuf<uint8_t> d_tmp;

    h = timer.begin("pileup_scan", 3);
    d_cover.zero();
    cover_diff(R, d_blank.p, d_cover.p, s);
    cover_scan(d_cover.p, L + 1, sp, s);
    timer.end(h);
    // K2 proper: one CTA per stripe of positions finds, buckets and merges the non-reference 3-mers in shared memory and
    // writes the finished Msa entries (np2_kernels.cu k_pileup_stripe).  Exact mode sizes the entry arrays from a
    // counting run of the same kernel.
    MsaDev m;
    m.L = L;
    m.cnt = cd.c;
    DBuf<uint32_t> d_sp_off, d_gcount, d_gfirst, d_gbesti, d_dense_cnt, d_dense_besti, d_n_emit, d_emit_off;
    DBuf<uint16_t> d_gbases, d_gdelta, d_sp_cnt;
    DBuf<int64_t> d_gscore, d_dense_score;
    DBuf<uint8_t> d_multi;
    d_sp_off.alloc(L + 1, s);
    d_sp_cnt.alloc(L + 1, s);
    d_dense_cnt.alloc(L, s);
    d_dense_besti.alloc(L, s);
    d_dense_score.alloc(L, s);
    d_multi.alloc(L, s);
    d_n_emit.alloc(L + 1, s);
    d_emit_off.alloc(L + 1, s);
    m.sp_off = d_sp_off.p;
    m.sp_cnt = d_sp_cnt.p;
    m.cover = d_cover.p;
    m.dense_cnt = d_dense_cnt.p;
    m.dense_besti = d_dense_besti.p;
    m.dense_score = d_dense_score.p;
    m.multi = d_multi.p;
    m.code = d_code.p;
    uint32_t G = spec ? caps.c[C_G] : 0;
    if (!spec) {
        h = timer.begin("pileup_count", 1);
        pileup_stripe(R, d_blank.p, d_code.p, d_odd_off.p, d_odd_list.p, m, 0, cd, d_n_emit.p, true, s);
        timer.end(h);
        G = cnt_get(C_G);
        counts_reset_pileup(cd, s);
    }
    d_gcount.alloc(std::max(G, 1u), s);
    d_gfirst.alloc(std::max(G, 1u), s);
    d_gbesti.alloc(std::max(G, 1u), s);
    d_gbases.alloc(std::max(G, 1u), s);
    d_gdelta.alloc(std::max(G, 1u), s);
    d_gscore.alloc(std::max(G, 1u), s);
    m.g_bases = d_gbases.p;
    m.g_delta = d_gdelta.p;
    m.g_count = d_gcount.p;
    m.g_first = d_gfirst.p;
    m.g_besti = d_gbesti.p;
    m.g_score = d_gscore.p;
    h = timer.begin("pileup_stripe", 1);
    if (dump_iter >= 0) d_dense_besti.zero();  // the stage getter reports besti of every position
    pileup_stripe(R, d_blank.p, d_code.p, d_odd_off.p, d_odd_list.p, m, G, cd, d_n_emit.p, false, s);
    timer.end(h);

    /* ---------------- K3: DP over runs, backtrack, consensus */
    DBuf<uint32_t> d_run_start;
    const uint32_t cap_runs_alloc = spec ? caps.c[C_NRUNS] : L / 2 + 2;  // runs are separated by single-entry positions
    d_run_start.alloc(cap_runs_alloc, s);
    h = timer.begin("dp_runs_select", 1);
    runs_select(d_multi.p, L, d_run_start.p, cap_runs_alloc, cd, sp, s);
    timer.end(h);
    const uint32_t n_runs = cnt_get(C_NRUNS);
    DBuf<uint32_t> d_long_runs;
    d_long_runs.alloc(std::max(n_runs, 1u), s);
    h = timer.begin("dp_runs", 2);
    dp_runs(m, d_run_start.p, n_runs, cd, d_long_runs.p, s);
    timer.end(h);
    // consensus bases: at most one per position plus the insertions the best path takes (<= the sparse entries)
    const uint32_t cap_n = spec ? caps.c[C_N] : (uint32_t)std::min<uint64_t>((uint64_t)L + G + 16, 0xFFFFFFF0ull);
    DBuf<uint32_t> &d_cpos = jd_cpos;
    DBuf<uint8_t> &d_cbase = jd_cbase, &d_cflags = jd_cflags;
    DBuf<uint32_t> d_events;
    d_cpos.alloc(cap_n, s);
    d_cbase.alloc(cap_n, s);
    d_cflags.alloc(cap_n, s);
    h = timer.begin("consensus_emit", 4);
    emit_count_runs(m, d_run_start.p, n_runs, cd, d_n_emit.p, s);
    emit_offsets(m, d_n_emit.p, d_emit_off.p, d_cpos.p, d_cbase.p, d_cflags.p, cap_n, cd, sp, s);
    emit_write(m, d_run_start.p, n_runs, cd, d_n_emit.p, d_emit_off.p, d_cpos.p, d_cbase.p, d_cflags.p, s);
    timer.end(h);
    const uint32_t N = cnt_get(C_N);
    d_events.alloc(std::max(N, 1u), s);
    h = timer.begin("consensus_emit", 1);
    // speculative mode: the event arrays below are sized from the remembered count, so THAT is the capacity the select
    // has to defend (more events than it => abort => the pass is repeated in exact mode), not the size of d_events
    events_select(d_cflags.p, N, d_events.p, spec ? std::min(std::max(N, 1u), caps.c[C_NEV]) : std::max(N, 1u), cd, sp, s);
    timer.end(h);
    const uint32_t n_ev = cnt_get(C_NEV);
    auto check_total = [&]() {
        if ((long long)hc->q[Q_TOTAL] < 0)
            throw np2::Error(NP2_ERR_UNSUPPORTED,
                             "best path has a negative total score (main.rs:1680 picks the default 3-mer): not supported");
    };
    auto check_perr = [&]() {
        const uint32_t e = hc->c[C_PERR];
        if (e == 4) throw np2::Error(NP2_ERR_INTERNAL, "read pair outside its index window");
        if (e == 5) throw np2::Error(NP2_ERR_UNSUPPORTED, "more than 1024 non-reference 3-mers at one position (coverage too deep)");
        if (e == 6) throw np2::Error(NP2_ERR_UNSUPPORTED, "more than 65535 copies or kinds of a 3-mer at one position");
    };
    if (!spec) {
        check_total();
        check_perr();
    }

    /* ---------------- LQ regions on the device (np2_regions.cu) */
    RegionDev rd;
    rd.cnt = cd.c;
    rd.events = d_events.p;
    rd.cflags = d_cflags.p;
    rd.cbase = d_cbase.p;
    rd.cpos = d_cpos.p;
    DBuf<uint32_t> d_ev_close, d_c_t, d_c_start, d_c_end, d_c_a, d_c_b, d_c_head, d_c_hrank;
    DBuf<uint8_t> d_ev_boundary, d_ev_closes;
    DBuf<uint32_t> d_rstart, d_rend, d_ra, d_rb;
    d_ev_close.alloc(std::max(n_ev, 1u), s);
    d_ev_boundary.alloc(std::max(n_ev, 1u), s);
    d_ev_closes.alloc(std::max(n_ev, 1u), s);
    d_c_t.alloc(std::max(n_ev, 1u), s);
    rd.ev_close = d_ev_close.p;
    rd.ev_boundary = d_ev_boundary.p;
    rd.ev_closes = d_ev_closes.p;
    rd.c_t = d_c_t.p;
    h = timer.begin("regions", 2);
    regions_event_close(rd, n_ev, s);
    regions_cand_select(rd, n_ev, spec ? std::min(std::max(n_ev, 1u), caps.c[C_NCAND]) : std::max(n_ev, 1u), cd, sp, s);
    timer.end(h);
    const uint32_t n_cand = cnt_get(C_NCAND);
    d_c_start.alloc(std::max(n_cand, 1u), s);
    d_c_end.alloc(std::max(n_cand, 1u), s);
    d_c_a.alloc(std::max(n_cand, 1u), s);
    d_c_b.alloc(std::max(n_cand, 1u), s);
    d_c_head.alloc(n_cand + 1, s);
    d_c_hrank.alloc(n_cand + 1, s);
    rd.c_start = d_c_start.p;
    rd.c_end = d_c_end.p;
    rd.c_a = d_c_a.p;
    rd.c_b = d_c_b.p;
    rd.c_head = d_c_head.p;
    rd.c_hrank = d_c_hrank.p;
    h = timer.begin("regions", 3);
    regions_make(rd, n_cand, s);
    regions_rank(rd, n_cand, spec ? caps.c[C_NREG] : 0xFFFFFFFFu, cd, sp, s);
    timer.end(h);
    uint32_t nreg = cnt_get(C_NREG);
    d_rstart.alloc(std::max(nreg, 1u), s);
    d_rend.alloc(std::max(nreg, 1u), s);
    d_ra.alloc(std::max(nreg, 1u), s);
    d_rb.alloc(std::max(nreg, 1u), s);
    rd.r_start = d_rstart.p;
    rd.r_end = d_rend.p;
    rd.r_a = d_ra.p;
    rd.r_b = d_rb.p;
    h = timer.begin("regions", 1);
    regions_out(rd, n_cand, s);
    timer.end(h);
    Regions rg;  // host copy only when needed (dump, final iteration)
    auto fetch_regions = [&]() {
        rg.start.resize(nreg);
        rg.end.resize(nreg);
        rg.a.resize(nreg);
        rg.b.resize(nreg);
        if (nreg) {
            d_rstart.download(rg.start.data(), nreg);
            d_rend.download(rg.end.data(), nreg);
            d_ra.download(rg.a.data(), nreg);
            d_rb.download(rg.b.data(), nreg);
            NP2_CUDA(cudaStreamSynchronize(s));
            d2h += (uint64_t)nreg * 16;
        }
    };

    auto dump_stage1 = [&]() {  // exact mode only
        // Msa in the reference's order: reference 3-mer first (p >= 2), then the sorted sparse ones
        std::vector<uint32_t> sp_off(L + 1), gc(G), gb(G), dc(L), db(L);
        std::vector<uint16_t> gba(G), gde(G), sp_cnt(L + 1);
        std::vector<uint8_t> code(L);
        d_sp_off.download(sp_off.data(), L);
        d_sp_cnt.download(sp_cnt.data(), L);
        d_gcount.download(gc.data(), G);
        d_gbesti.download(gb.data(), G);
        d_gbases.download(gba.data(), G);
        d_gdelta.download(gde.data(), G);
        d_dense_cnt.download(dc.data(), L);
        d_dense_besti.download(db.data(), L);
        d_code.download(code.data(), L);
        dm_dp_pos.resize(N);
        dm_dp_base.resize(N);
        dm_dp_flags.resize(N);
        d_cpos.download(dm_dp_pos.data(), N);
        d_cbase.download(dm_dp_base.data(), N);
        d_cflags.download(dm_dp_flags.data(), N);
        NP2_CUDA(cudaStreamSynchronize(s));
        dm_msa_off.assign(1, 0);
        for (uint32_t p = 0; p < L; p++) {
            if (p >= 2) {
                dm_msa_bases.push_back((uint16_t)(code[p - 2] << 8 | code[p - 1] << 4 | code[p]));
                dm_msa_delta.push_back(0);
                dm_msa_count.push_back(dc[p]);
                dm_msa_besti.push_back(db[p]);
            }
            for (uint32_t g = sp_off[p]; g < sp_off[p] + sp_cnt[p]; g++) {
                dm_msa_bases.push_back(gba[g]);
                dm_msa_delta.push_back(gde[g]);
                dm_msa_count.push_back(gc[g]);
                dm_msa_besti.push_back(gb[g]);
            }
            dm_msa_off.push_back(dm_msa_bases.size());
        }
        fetch_regions();
        dm_reg_start = rg.start;
        dm_reg_end = rg.end;
    };
    uint32_t edge_pos[2] = {0, 0};  // ConsensusBase.pos of the first / last DP base (FASTA header)
    auto fetch_edge_pos = [&](uint32_t n_true) {
        if (!n_true) return;
        uint32_t *pe = reinterpret_cast<uint32_t *>(hc + 1);  // two pinned words behind the counts
        NP2_CUDA(cudaMemcpyAsync(pe, d_cpos.p, 4, cudaMemcpyDeviceToHost, s));
        NP2_CUDA(cudaMemcpyAsync(pe + 1, d_cpos.p + (n_true - 1), 4, cudaMemcpyDeviceToHost, s));
        NP2_CUDA(cudaStreamSynchronize(s));
        n_sync++;
        edge_pos[0] = pe[0];
        edge_pos[1] = pe[1];
    };
    auto note_sizes = [&]() {  // after a fetch: what bench.py and the next pass's capacities want to know
        stats[0] = hc->c[C_NREC];
        stats[1] = hc->c[C_G];
        stats[2] = hc->c[C_NRUNS];
        stats[3] = hc->c[C_N];
        stats[4] = hc->c[C_NREG];
        stats[5] = hc->c[C_NPAIRS];
        res_N = hc->c[C_N];
    };
    auto remember = [&]() {
        // the counts of THIS pass (its fetches only ever add information), not the largest ever seen: a context that
        // polishes contigs of shrinking size would otherwise keep the capacities of its largest one, scaled only by the
        // length ratio to the previous contig, and do several times the work on the small ones
        if (!hint.in_pass) {
            hint.h = CountsHost();
            hint.in_pass = true;
        }
        hint.valid = true;
        hint.L = L;
        hint.cols = ing.total_cols;
        for (int i = 0; i < C_COUNT; i++) hint.h.c[i] = std::max(hint.h.c[i], hc->c[i]);
        for (int i = 0; i < Q_COUNT; i++) hint.h.q[i] = std::max(hint.h.q[i], hc->q[i]);
        hint.h.c[C_ABORT] = 0;
        hint.h.q[Q_TOTAL] = hint.h.q[Q_SHIFT] = 0;
    };
    stats[7]++;
    uint32_t iter = iter0;
    if (!spec && nreg == 0) {  // main.rs:1638-1640: no LQ region; nothing can be dropped, the DP consensus is the answer
        note_sizes();
        hint.h = CountsHost();
        remember();
        for (;; iter++) {
            if ((int32_t)iter == dump_iter) dump_stage1();
            if (iter + 1 < opt.iter_count) continue;
            fetch_edge_pos(N);
            res_patch.reset(0);
            res_base.resize(std::max(N, 1u));
            res_base.n = N;
            d_cbase.download(res_base.p, N);
            NP2_CUDA(cudaStreamSynchronize(s));
            n_sync++;
            d2h += N;
            res_first = edge_pos[0];
            res_last = edge_pos[1];
            res_pos_valid = false;
            return iter + 1;
        }
    }

    /* ---------------- K4: which reads cover which region, candidates, kscore — all on the device */
    np2_table *t0 = tables[0];
    const uint32_t k0 = t0->dev.k;
    if (k0 >= 32) throw np2::Error(NP2_ERR_UNSUPPORTED, "the smallest yak table must have k < 32 (main.rs:1432-1434)");
    GenoDev g;
    g.cnt = cd.c;
    DBuf<uint32_t> d_rd_s, d_rd_j, d_rd_np, d_rd_poff;
    d_rd_s.alloc(n_reads + 1, s);
    d_rd_j.alloc(n_reads + 1, s);
    d_rd_np.alloc(n_reads + 1, s);
    d_rd_poff.alloc(n_reads + 1, s);
    g.start = d_rstart.p;
    g.end = d_rend.p;
    g.rd_s = d_rd_s.p;
    g.rd_j = d_rd_j.p;
    g.rd_np = d_rd_np.p;
    g.rd_poff = d_rd_poff.p;
    g.rd_order = d_order.p;
    h = timer.begin("read_ranges", 4);
    geno_read_cursor(g, R, d_blank.p, s);
    geno_cursor_min(d_rd_s.p, n_reads, cd.c + C_ABORT, sp, s);  // the cursor only ever moves down (main.rs:1446-1448)
    geno_read_ranges(g, R, d_blank.p, k0, s);
    geno_pair_offsets(g, n_reads, spec ? caps.c[C_NPAIRS] : 0xFFFFFFFFu, cd, sp, s);
    timer.end(h);
    const uint32_t n_pairs = cnt_get(C_NPAIRS);

    const uint64_t nslot = (uint64_t)std::max(nreg, 1u) * kMaxCand;
    DBuf<uint32_t> d_p_len, d_c_src, d_c_len, d_c_order, d_r_ncand, d_r_bytes, d_r_nedge, d_r_seed_len, d_r_nsurv;
    DBuf<uint64_t> d_p_kmer, d_c_kmer, d_c_off, d_r_pool_off, d_r_edge_off, d_r_seed_off;
    DBuf<uint16_t> d_c_ks;
    DBuf<uint8_t> d_c_rep, d_r_lable, d_r_surv, d_gpool;
    d_p_len.alloc(std::max(n_pairs, 1u), s);
    d_p_kmer.alloc(std::max(n_pairs, 1u), s);
    d_c_src.alloc(nslot, s);
    d_c_len.alloc(nslot, s);
    d_c_order.alloc(nslot, s);
    d_c_kmer.alloc(nslot, s);
    d_c_off.alloc(nslot, s);
    d_c_ks.alloc(nslot, s);
    d_c_rep.alloc(nslot, s);
    d_r_ncand.alloc(nreg + 1, s);
    d_r_bytes.alloc(nreg + 1, s);
    d_r_nedge.alloc(nreg + 1, s);
    d_r_seed_len.alloc(nreg + 1, s);
    d_r_nsurv.alloc(nreg + 1, s);
    d_r_pool_off.alloc(nreg + 1, s);
    d_r_edge_off.alloc(nreg + 1, s);
    d_r_seed_off.alloc(nreg + 1, s);
    d_r_lable.alloc(nreg + 1, s);
    d_r_surv.alloc(nslot, s);
    DBuf<uint32_t> d_long_list;
    d_long_list.alloc(nslot, s);
    g.long_list = d_long_list.p;
    g.p_len = d_p_len.p;
    g.p_kmer = d_p_kmer.p;
    g.c_src = d_c_src.p;
    g.c_len = d_c_len.p;
    g.c_order = d_c_order.p;
    g.c_kmer = d_c_kmer.p;
    g.c_off = d_c_off.p;
    g.c_kscore = d_c_ks.p;
    g.c_rep = d_c_rep.p;
    g.r_ncand = d_r_ncand.p;
    g.r_bytes = d_r_bytes.p;
    g.r_nedge = d_r_nedge.p;
    g.r_seed_len = d_r_seed_len.p;
    g.r_nsurv = d_r_nsurv.p;
    g.r_pool_off = d_r_pool_off.p;
    g.r_edge_off = d_r_edge_off.p;
    g.r_seed_off = d_r_seed_off.p;
    g.r_lable = d_r_lable.p;
    g.r_surv = d_r_surv.p;

    h = timer.begin("cand_scan", 1);
    geno_pair_scan(g, R, k0, n_pairs, s);
    timer.end(h);
    h = timer.begin("region_select", 2);
    geno_region_select(g, R, d_blank.p, d_code.p, L, k0, max_span, d_first_ge.p, pileup_stripe_width(), nreg, s);
    geno_pool_offsets(g, nreg, spec ? caps.q[Q_POOL] : ~0ULL, cd, sp, s);
    timer.end(h);
    uint64_t pool_bytes = caps.q[Q_POOL];
    if (!spec) {
        fetch_counts();
        pool_bytes = hc->q[Q_POOL];
    }
    d_gpool.alloc(pool_bytes + 16, s);
    g.pool = d_gpool.p;
    h = timer.begin("cand_write", 1);
    geno_cand_write(g, R, d_code.p, L, k0, nreg, s);
    timer.end(h);
    h = timer.begin("yak_probe", 2);
    geno_cand_kscore(g, t0->dev, opt.min_kmer_count, nreg, cd, s);
    timer.end(h);

    auto dump_candidates = [&](bool with_lable) {  // exact mode only
        std::vector<uint32_t> ncand(nreg), clen(nslot), cord(nslot);
        std::vector<uint64_t> ckm(nslot), coff(nslot);
        std::vector<uint16_t> cks(nslot);
        std::vector<uint8_t> pool(pool_bytes + 16), lab(nreg);
        d_r_ncand.download(ncand.data(), nreg);
        d_c_len.download(clen.data(), nslot);
        d_c_order.download(cord.data(), nslot);
        d_c_kmer.download(ckm.data(), nslot);
        d_c_off.download(coff.data(), nslot);
        d_c_ks.download(cks.data(), nslot);
        d_gpool.download(pool.data(), pool_bytes);
        if (with_lable) d_r_lable.download(lab.data(), nreg);
        NP2_CUDA(cudaStreamSynchronize(s));
        if (with_lable) {
            dm_reg_lable = lab;
            return;
        }
        dm_can_roff.assign(1, 0);
        dm_can_seq_off.assign(1, 0);
        for (uint32_t r = 0; r < nreg; r++) {
            for (uint32_t c = 0; c < ncand[r]; c++) {
                const uint64_t sl = (uint64_t)r * kMaxCand + c;
                dm_can_order.push_back(cord[sl]);
                dm_can_kscore.push_back(cks[sl]);
                dm_can_kmer.push_back(ckm[sl]);
                dm_can_seq.insert(dm_can_seq.end(), pool.data() + coff[sl], pool.data() + coff[sl] + clen[sl]);
                dm_can_seq_off.push_back(dm_can_seq.size());
            }
            dm_can_roff.push_back(dm_can_order.size());
        }
    };
    // mark_hete_lqseqs zeroes kscores in place: keep the retrieve_kmer_count values for a re-used iteration
    DBuf<uint16_t> d_c_ks_orig;
    const bool may_reuse = iter0 + 2 <= opt.iter_count;  // a non-final iteration may be followed by a re-used one
    if (may_reuse) {
        d_c_ks_orig.alloc(nslot, s);
        NP2_CUDA(cudaMemcpyAsync(d_c_ks_orig.p, d_c_ks.p, nslot * 2, cudaMemcpyDeviceToDevice, s));
    }
    const uint32_t n_ids0 = (uint32_t)as_read.size();
    bool unchecked_edges = false;  // a non-final iteration of this stretch went on without reading its counts back
  for (;; iter++) {
    const bool final_iter = iter + 1 == opt.iter_count;
    const bool dump = (int32_t)iter == dump_iter;
    if (iter > iter0 && may_reuse) NP2_CUDA(cudaMemcpyAsync(d_c_ks.p, d_c_ks_orig.p, nslot * 2, cudaMemcpyDeviceToDevice, s));
    if (dump) dump_stage1();
    if (dump) dump_candidates(false);

    if (!final_iter) {
        /* ---------------- heterozygous regions, agreement edges (device); Louvain (host) — main.rs:1544-1552 */
        h = timer.begin("region_hete", 2);
        geno_region_hete(g, nreg, s);
        geno_edge_offsets(g, nreg, cd, sp, s);
        timer.end(h);
        // dense pair accumulator (np2_geno.cu k_edges_accum): one slot per (x, y) inside x's index window; the
        // number of slots only depends on the reads (np2_job::run)
        if (pair_slots >= (1ull << 31)) throw np2::Error(NP2_ERR_UNSUPPORTED, "more than 2^31 overlapping read pairs");
        const uint32_t n_slots = (uint32_t)pair_slots;
        uint32_t id_bits = 1;  // read orders are < as_read.size()
        while ((1ull << id_bits) < as_read.size()) id_bits++;
        // speculative mode: the pair kernels are only enqueued when the last pass of this kind had heterozygous
        // regions; a pass that finds some although the last one had none is repeated
        uint64_t n_edges = hint.h.q[Q_EDGES];
        const bool edges_enqueued = n_edges != 0;
        if (!spec) {
            fetch_counts();
            n_edges = hc->q[Q_EDGES];
        }
        if (dump) dump_candidates(true);
        std::vector<uint32_t> drop;
        DBuf<unsigned long long> d_acc;
        DBuf<uint8_t> d_flags;  // has | bad_v | in_ref
        DBuf<float> d_refw, d_aw;
        DBuf<uint32_t> d_aoff, d_ato;
        uint32_t nu = 0, n_dir_cap = 0;
        if (n_edges) {
            d_acc.alloc(std::max<uint64_t>(n_slots, 1), s);
            h = timer.begin("pair_edges", 2);
            d_acc.zero();
            geno_edges_accum(g, d_pair_off.p, d_acc.p, nreg, cd, s);
            timer.end(h);
            // level 0 of the phasing graph on the device, straight from the accumulator (np2_geno.cu k_phase_adj): per-read
            // flags + CSR adjacency with every neighbour list in ascending order
            const bool asref = opt.model == 0, use_all = opt.use_all_reads != 0;
            DBuf<uint32_t> d_deg;
            d_flags.alloc((size_t)n_ids0 * 3, s);
            d_refw.alloc(n_ids0, s);
            d_deg.alloc(n_ids0 + 1, s);
            d_aoff.alloc(n_ids0 + 1, s);
            PhaseDev pd;
            pd.has = d_flags.p;
            pd.bad_v = d_flags.p + n_ids0;
            pd.in_ref = d_flags.p + 2 * (size_t)n_ids0;
            pd.ref_w = d_refw.p;
            h = timer.begin("phase_graph", 6);
            d_flags.zero();
            d_refw.zero();
            phase_ref_acc(d_acc.p, d_pair_off.p, n_ids0, cd, pd, asref, use_all, s);
            phase_adj_count(d_acc.p, d_pair_off.p, d_as_pos.p, n_ids0, max_span, cd, pd, use_all, d_deg.p, s);
            const uint32_t cap_dir = spec ? caps.c[C_NDIR] : 0xFFFFFFFFu;
            phase_adj_offsets(d_deg.p, d_aoff.p, n_ids0, cap_dir, cd, sp, s);
            timer.end(h);
            n_dir_cap = spec ? cap_dir : cnt_get(C_NDIR);
            if (!spec) check_perr();
            nu = 1;  // "the pair kernels were enqueued"
            d_ato.alloc(std::max(n_dir_cap, 1u), s);
            d_aw.alloc(std::max(n_dir_cap, 1u), s);
            h = timer.begin("phase_graph", 1);
            phase_adj_fill(d_acc.p, d_pair_off.p, d_as_pos.p, n_ids0, max_span, cd, pd, use_all, d_aoff.p, d_ato.p, d_aw.p,
                           n_dir_cap, s);
            timer.end(h);
        }
        // the reduced pair records in (x, y) order (key = x << 32 | y): only the stage seam and the general phasing path
        // want them
        auto pair_records = [&](std::vector<uint64_t> &keys, std::vector<long long> &vals, uint32_t nu_exact) {
            keys.resize(nu_exact);
            vals.resize(nu_exact);
            if (!nu_exact) return;
            DBuf<uint32_t> d_sel;
            DBuf<uint64_t> d_uk;
            DBuf<long long> d_uv;
            d_sel.alloc(nu_exact, s);
            d_uk.alloc(nu_exact, s);
            d_uv.alloc(nu_exact, s);
            geno_edges_select(d_acc.p, n_slots, d_sel.p, nu_exact, cd, sp, s);  // rewrites C_NU with the same value
            geno_edges_finish(d_sel.p, nu_exact, d_pair_off.p, n_ids0, d_acc.p, d_uk.p, d_uv.p, cd, s);
            d_uk.download(keys.data(), nu_exact);
            NP2_CUDA(cudaMemcpyAsync(vals.data(), d_uv.p, (size_t)nu_exact * 8, cudaMemcpyDeviceToHost, s));
            NP2_CUDA(cudaStreamSynchronize(s));
            d2h += (uint64_t)nu_exact * 16;
        };
        if (dump && n_edges) {  // stage seam: the pair weights as they leave K6 (exact mode)
            fetch_counts();
            std::vector<long long> v;
            pair_records(dm_pair_key, v, hc->c[C_NU]);
            dm_pair_val.assign(v.begin(), v.end());
        }
        const bool was_spec = spec;
        if (spec && !edges_enqueued && !dump) {
            // No pair kernel was enqueued (the last pass of this kind had no heterozygous region).  If this one has some,
            // the pass is repeated in exact mode anyway; if it has none, no read can be dropped and the next iteration
            // is this one again: nothing to wait for here.  The stretch goes on into the next iteration, and the
            // heterozygous-region count is looked at where the stretch does end.
            unchecked_edges = true;
            continue;
        }
        if (spec) {  // the one synchronisation of the speculative stretch
            timer.hbegin();
            segment_end();
            timer.hend("host:segment_sync");
            check_total();
            if (hc->c[C_NREG] == 0) throw Respeculate();  // the no-region path is only taken in exact mode
            if (hc->q[Q_EDGES] != 0 && !edges_enqueued) throw Respeculate();
            check_perr();
            nreg = hc->c[C_NREG];
        }
        note_sizes();
        stats[6] = hc->c[C_NU];
        remember();
        const uint32_t nu_true = hc->c[C_NU];
        if (n_edges && nu && nu_true) {
            const bool asref = opt.model == 0, use_all = opt.use_all_reads != 0;
            const uint32_t n_dir = hc->c[C_NDIR];  // exact: the counts were fetched above
            (void)was_spec;
            // one pinned block: aoff | ato | aw | ref_w | flags
            const size_t o_ato = ((size_t)(n_ids0 + 1) * 4 + 15) & ~(size_t)15, o_aw = o_ato + (((size_t)n_dir * 4 + 15) & ~(size_t)15);
            const size_t o_rw = o_aw + (((size_t)n_dir * 4 + 15) & ~(size_t)15), o_fl = o_rw + (((size_t)n_ids0 * 4 + 15) & ~(size_t)15);
            sc->p_phase.resize(o_fl + (size_t)n_ids0 * 3 + 16);
            uint8_t *pb = sc->p_phase.p;
            d_aoff.download(reinterpret_cast<uint32_t *>(pb), n_ids0 + 1);
            d_ato.download(reinterpret_cast<uint32_t *>(pb + o_ato), n_dir);
            d_aw.download(reinterpret_cast<float *>(pb + o_aw), n_dir);
            d_refw.download(reinterpret_cast<float *>(pb + o_rw), n_ids0);
            d_flags.download(pb + o_fl, (size_t)n_ids0 * 3);
            NP2_CUDA(cudaStreamSynchronize(s));
            n_sync++;
            d2h += (uint64_t)n_dir * 8 + (uint64_t)n_ids0 * 11;
            timer.hbegin();
            drop = phase_reads_csr(n_ids0, reinterpret_cast<const uint32_t *>(pb), reinterpret_cast<const uint32_t *>(pb + o_ato),
                                   reinterpret_cast<const float *>(pb + o_aw), pb + o_fl, pb + o_fl + n_ids0,
                                   pb + o_fl + 2 * (size_t)n_ids0, reinterpret_cast<const float *>(pb + o_rw), asref, [&]() {
                                       // a community has to be declustered: the general path wants the pair records
                                       std::vector<uint64_t> ukeys;
                                       std::vector<long long> uvals;
                                       pair_records(ukeys, uvals, nu_true);
                                       return phase_reads_general(ukeys.data(), uvals.data(), nu_true, asref, use_all);
                                   });
            timer.hend("host:phase_reads");
            static const char *kPhase[4] = {"host:phase_reads.build", "host:phase_reads.move", "host:phase_reads.aggregate",
                                            "host:phase_reads.communities"};
            for (int x = 0; x < 4; x++) timer.ms[timer.id(kPhase[x])] += np2::phase_last_ms()[x];
        }
        for (uint32_t a : drop) {
            if (a == 0 || a >= as_read.size()) throw np2::Error(NP2_ERR_INTERNAL, "phasing returned a bad read index");
            h_blank[as_read[a]] = 1;
            dm_dropped.push_back(a);
        }
        if (drop.empty()) continue;  // same reads => the next iteration is this one again
        Uploader{sc->p_h2d, sc->h2d_used, s}.put(d_blank.p, h_blank.data(), n_reads);
        return iter + 1;
    }

    /* ---------------- final: seed alleles (device), then re-check with every table (main.rs:1527-1543) */
    h = timer.begin("region_seed", 1);
    geno_region_seed(g, opt.max_indel_len, nreg, cd, iter > iter0, s);
    timer.end(h);
    /* ---- what the host needs for the re-check: regions, every region's seed string, the survivors of the regions
     *      that stay RECH, and the DP bases (flanks).  Everything else stays on the device. */
    AssembleDev ad;
    ad.cnt = cd.c;
    ad.cbase = d_cbase.p;
    ad.pool = d_gpool.p;
    ad.r_a = d_ra.p;
    ad.r_b = d_rb.p;
    ad.r_seed_len = d_r_seed_len.p;
    ad.r_seed_off = d_r_seed_off.p;
    DBuf<long long> d_q_delta, d_q_shift;
    DBuf<uint32_t> d_q_seedlen, d_rech_bytes, d_ent_off;
    DBuf<uint64_t> d_q_seedoff, d_rech_boff;
    d_q_delta.alloc(nreg + 1, s);
    d_q_shift.alloc(nreg + 1, s);
    d_q_seedlen.alloc(nreg + 1, s);
    d_q_seedoff.alloc(nreg + 1, s);
    d_rech_bytes.alloc(nreg + 1, s);
    d_rech_boff.alloc(nreg + 1, s);
    d_ent_off.alloc(nreg + 1, s);
    ad.q_delta = d_q_delta.p;
    ad.q_shift = d_q_shift.p;
    ad.q_seedlen = d_q_seedlen.p;
    ad.q_seedoff = d_q_seedoff.p;
    Patched &pc = res_patch;
    uint32_t *seed_len = nullptr;   // full path: staged copies of the per-region seed (r order)
    uint64_t *seed_off = nullptr;
    auto phase1_sizes = [&]() {
        h = timer.begin("seed_gather", 5);
        assemble_sizes(ad, nreg, s);
        region_scan_u64(d_q_seedlen.p, d_q_seedoff.p, nreg, Q_SEEDS, cd, sp, s);
        rech_sizes(g, d_rech_bytes.p, nreg, s);
        region_scan_u64(d_rech_bytes.p, d_rech_boff.p, nreg, Q_RECH, cd, sp, s);
        region_scan_u32(d_r_nsurv.p, d_ent_off.p, nreg, C_NENT, cd, sp, s);
        timer.end(h);
    };
    auto raise_gerr = [&]() {
        const uint32_t gerr = hc->c[C_GERR];
        if (gerr == 1) throw np2::Error(NP2_ERR_FORMAT, "LQ region without any candidate (reference would panic)");
        if (gerr == 2) throw np2::Error(NP2_ERR_FORMAT, "the first lqseq is not ref.");
        if (gerr == 3) throw np2::Error(NP2_ERR_FORMAT, "no candidate survives retain_sort_seqs (reference would panic)");
    };
    /* Sparse host view (the normal case): the host only looks at the RECH regions and at what lies within
     * kRecheckWindow DP bases of them, plus the first and the last region (FASTA header span).  Everything is selected,
     * compacted and gathered on the device; per-region work on the host is O(selected), not O(nreg).  Returns false
     * (nothing committed) when the windows would cover most of the contig or the layout check fails. */
    std::vector<uint32_t> sv_sub;         // selected regions, ascending r
    std::vector<uint64_t> sv_init_off;    // per selected region (q' order): the seed it started with
    std::vector<uint32_t> sv_init_len;
    long long sv_shift0 = 0;
    bool sparse_view = false;
    uint32_t N_true = N;
    auto build_sparse_view = [&]() -> bool {
        DBuf<uint8_t> d_near;
        DBuf<uint32_t> d_sub, d_win_lo, d_win_len;
        DBuf<uint64_t> d_win_off;
        d_near.alloc(nreg + 1, s);
        d_sub.alloc(nreg + 1, s);
        d_win_lo.alloc(nreg + 1, s);
        d_win_len.alloc(nreg + 1, s);
        d_win_off.alloc(nreg + 1, s);
        SubMeta sm;
        DBuf<uint32_t> m_start, m_end, m_a, m_b, m_seed_len, m_nsurv, m_ent_off;
        DBuf<uint64_t> m_seed_off, m_qseedoff;
        DBuf<uint8_t> m_lable;
        for (DBuf<uint32_t> *d : {&m_start, &m_end, &m_a, &m_b, &m_seed_len, &m_nsurv, &m_ent_off}) d->alloc(nreg + 1, s);
        m_seed_off.alloc(nreg + 1, s);
        m_qseedoff.alloc(nreg + 1, s);
        m_lable.alloc(nreg + 1, s);
        sm.start = m_start.p, sm.end = m_end.p, sm.a = m_a.p, sm.b = m_b.p, sm.seed_len = m_seed_len.p;
        sm.nsurv = m_nsurv.p, sm.ent_off = m_ent_off.p, sm.seed_off = m_seed_off.p, sm.q_seedoff = m_qseedoff.p;
        sm.lable = m_lable.p;
        int hh = timer.begin("seed_gather", 2);
        d_near.zero();
        near_mark(cd.c, nreg, d_r_lable.p, d_ra.p, d_rb.p, d_near.p, s);
        ad.near = d_near.p;
        timer.end(hh);
        phase1_sizes();
        hh = timer.begin("seed_gather", 5);
        region_scan_i64(d_q_delta.p, d_q_shift.p, nreg, Q_SHIFT, cd, sp, s);
        window_sizes(cd.c, nreg, d_r_lable.p, d_ra.p, d_rb.p, d_win_lo.p, d_win_len.p, s);
        region_scan_u64(d_win_len.p, d_win_off.p, nreg, Q_WIN, cd, sp, s);
        near_select(d_near.p, nreg, d_sub.p, cd, sp, s);
        sub_meta_gather(d_sub.p, cd.c, nreg, d_rstart.p, d_rend.p, d_ra.p, d_rb.p, d_r_lable.p, d_r_seed_len.p,
                        d_r_seed_off.p, d_r_nsurv.p, d_ent_off.p, d_q_seedoff.p, sm, s);
        timer.end(hh);
        timer.hbegin();
        if (spec) {
            segment_end();
            check_total();
            check_perr();
            if (hc->c[C_NREG] == 0) throw Respeculate();
            if (unchecked_edges && hc->q[Q_EDGES] != 0) throw Respeculate();
        } else {
            fetch_counts();
        }
        timer.hend("host:seed_sync1");
        note_sizes();
        remember();
        nreg = hc->c[C_NREG];
        N_true = hc->c[C_N];
        raise_gerr();
        const uint32_t nsub = hc->c[C_NSUB], n_ent = hc->c[C_NENT];
        const uint64_t seeds_bytes = hc->q[Q_SEEDS], rech_bytes = hc->q[Q_RECH], win_bytes = hc->q[Q_WIN];
        sv_shift0 = (long long)hc->q[Q_SHIFT];
        if (win_bytes > N_true / 2) return false;
        DBuf<uint8_t> d_seeds, d_rech_pool, d_win;
        DBuf<uint32_t> d_ent_order, d_ent_len;
        DBuf<uint64_t> d_ent_poff;
        d_seeds.alloc(seeds_bytes + 1, s);
        d_rech_pool.alloc(rech_bytes + 1, s);
        d_win.alloc(win_bytes + 1, s);
        d_ent_order.alloc(std::max(n_ent, 1u), s);
        d_ent_len.alloc(std::max(n_ent, 1u), s);
        d_ent_poff.alloc(std::max(n_ent, 1u), s);
        hh = timer.begin("seed_gather", 3);
        assemble_seed_gather(ad, d_seeds.p, nreg, s);
        if (n_ent) rech_gather(g, d_ent_off.p, d_rech_boff.p, d_ent_order.p, d_ent_len.p, d_ent_poff.p, d_rech_pool.p, nreg, s);
        if (win_bytes) gather_ranges(d_cbase.p, d_win_lo.p, d_win_off.p, cd.c + C_NREG, nreg, d_win.p, s);
        timer.end(hh);
        Stager st2(sc->p_stage2, s, (size_t)nsub * 56 + (size_t)n_ent * 16 + 4096);
        const uint32_t *h_sub = st2.fetch(d_sub.p, nsub);
        const uint32_t *ms = st2.fetch(m_start.p, nsub), *me = st2.fetch(m_end.p, nsub), *ma = st2.fetch(m_a.p, nsub);
        const uint32_t *mb = st2.fetch(m_b.p, nsub), *msl = st2.fetch(m_seed_len.p, nsub);
        const uint32_t *mns = st2.fetch(m_nsurv.p, nsub), *meo = st2.fetch(m_ent_off.p, nsub);
        const uint64_t *mso = st2.fetch(m_seed_off.p, nsub), *mqo = st2.fetch(m_qseedoff.p, nsub);
        const uint8_t *ml = st2.fetch(m_lable.p, nsub);
        const uint32_t *ent_order = st2.fetch(d_ent_order.p, n_ent), *ent_len = st2.fetch(d_ent_len.p, n_ent);
        const uint64_t *ent_poff = st2.fetch(d_ent_poff.p, n_ent);
        h_seeds.resize(seeds_bytes + 16);
        h_rech_pool.resize(rech_bytes + 16);
        h_win.resize(win_bytes + 16);
        d_seeds.download(h_seeds.data(), seeds_bytes);
        if (rech_bytes) d_rech_pool.download(h_rech_pool.data(), rech_bytes);
        if (win_bytes) d_win.download(h_win.data(), win_bytes);
        if (N_true) {  // ConsensusBase.pos of the first / last DP base (FASTA header), in the same round trip
            uint32_t *pe = reinterpret_cast<uint32_t *>(hc + 1);
            NP2_CUDA(cudaMemcpyAsync(pe, d_cpos.p, 4, cudaMemcpyDeviceToHost, s));
            NP2_CUDA(cudaMemcpyAsync(pe + 1, d_cpos.p + (N_true - 1), 4, cudaMemcpyDeviceToHost, s));
        }
        NP2_CUDA(cudaStreamSynchronize(s));
        n_sync++;
        if (N_true) {
            edge_pos[0] = reinterpret_cast<uint32_t *>(hc + 1)[0];
            edge_pos[1] = reinterpret_cast<uint32_t *>(hc + 1)[1];
        }
        timer.hend("host:seed_sync2");
        d2h += seeds_bytes + rech_bytes + win_bytes + (uint64_t)nsub * 53 + (uint64_t)n_ent * 16 + 64;
        // the view over the selected regions, ascending position (q' = nsub - 1 - i)
        pc.reset(nsub);
        pc.cbase = p_cbase.p;
        pc.N = N_true;
        sv_sub.assign(h_sub, h_sub + nsub);
        sv_init_off.resize(nsub);
        sv_init_len.resize(nsub);
        uint64_t rb = 0;
        for (uint32_t i = 0; i < nsub; i++) {
            const uint32_t q = nsub - 1 - i;
            pc.start[q] = ms[i];
            pc.end[q] = me[i];
            pc.a[q] = ma[i];
            pc.b[q] = mb[i];
            pc.lable[q] = ml[i];
            pc.seed[q].s = h_seeds.data() + mqo[i];
            pc.seed[q].len = msl[i];
            pc.seed[q].dev_off = mso[i];
            sv_init_off[q] = mso[i];
            sv_init_len[q] = msl[i];
            for (uint32_t x = 0; x < mns[i]; x++) {  // rech pool is laid out in r order, survivors in rank order
                Allele al;
                al.s = h_rech_pool.data() + rb;
                al.len = ent_len[meo[i] + x];
                al.order = ent_order[meo[i] + x];
                al.dev_off = ent_poff[meo[i] + x];
                rb += al.len;
                pc.cand[q].push_back(al);
            }
        }
        // windows into their true places + the layout check (DESIGN.md "Re-check windows")
        const uint32_t need = tables.back()->dev.k - 1;
        uint64_t woff = 0, prev_whi = 0;
        int64_t prev_rech = -1;
        bool ok = true;
        for (uint32_t q = 0; q < nsub && ok; q++) {
            if (!(pc.lable[q] & LABLE_RECH)) continue;
            const uint64_t wlo = pc.a[q] > kRecheckWindow ? pc.a[q] - kRecheckWindow : 0;
            const uint64_t whi = std::min<uint64_t>((uint64_t)pc.b[q] + kRecheckWindow, N_true);
            memcpy(p_cbase.p + wlo, h_win.data() + woff, whi - wlo);
            woff += whi - wlo;
            uint32_t got = 0, rr = q;  // left flank
            uint64_t i = pc.a[q];
            for (;;) {
                const uint64_t lo = rr > 0 ? pc.b[rr - 1] : 0;
                const uint64_t take = std::min<uint64_t>(need - got, i - lo);
                if (i - take < wlo) ok = false;
                got += (uint32_t)take;
                if (got >= need || rr == 0) break;
                rr--;
                i = pc.a[rr];
            }
            got = 0, rr = q, i = pc.b[q];  // right flank
            for (;;) {
                const uint64_t hi = rr + 1 < nsub ? pc.a[rr + 1] : N_true;
                const uint64_t take = std::min<uint64_t>(need - got, hi - i);
                if (i + take > whi) ok = false;
                got += (uint32_t)take;
                if (got >= need || rr + 1 >= nsub) break;
                rr++;
                i = pc.b[rr];
            }
            if (prev_rech >= 0 && pc.start[q] < pc.end[prev_rech] + need + 1 && pc.a[q] > prev_whi && wlo > prev_whi) ok = false;
            prev_rech = q;
            prev_whi = whi;
        }
        if (woff != win_bytes) ok = false;
        return ok;
    };
    if (!dump) {
        p_cbase.resize(std::max<uint64_t>(N, 1u));
        sparse_view = build_sparse_view();
        ad.near = nullptr;
    }
    if (!sparse_view) {
        phase1_sizes();
        timer.hbegin();
        if (spec) {
            segment_end();
            check_total();
            check_perr();
            if (hc->c[C_NREG] == 0) throw Respeculate();
            if (unchecked_edges && hc->q[Q_EDGES] != 0) throw Respeculate();
        } else {
            fetch_counts();
        }
        note_sizes();
        remember();
        nreg = hc->c[C_NREG];
        N_true = hc->c[C_N];
        raise_gerr();
        fetch_edge_pos(N_true);
        // round 1: everything whose size is known (one synchronisation, pinned destinations)
        p_cbase.resize(std::max(N_true, 1u));  // filled sparsely below: only the windows around RECH regions cross PCIe
        Stager st1(sc->p_stage, s, (size_t)nreg * 64 + 4096);
        const uint8_t *lab = st1.fetch(d_r_lable.p, nreg);
        seed_len = st1.fetch(d_r_seed_len.p, nreg);
        seed_off = st1.fetch(d_r_seed_off.p, nreg);
        const uint32_t *nsurv = st1.fetch(d_r_nsurv.p, nreg);
        const uint32_t *ent_off = st1.fetch(d_ent_off.p, nreg);
        const uint64_t *q_seedoff = st1.fetch(d_q_seedoff.p, nreg + 1);
        rg.start.resize(nreg);
        rg.end.resize(nreg);
        rg.a.resize(nreg);
        rg.b.resize(nreg);
        const uint32_t *h_rs = st1.fetch(d_rstart.p, nreg), *h_re = st1.fetch(d_rend.p, nreg);
        const uint32_t *h_ra = st1.fetch(d_ra.p, nreg), *h_rb = st1.fetch(d_rb.p, nreg);
        NP2_CUDA(cudaStreamSynchronize(s));
        n_sync++;
        timer.hend("host:seed_sync1");
        const uint64_t rech_bytes = hc->q[Q_RECH];
        memcpy(rg.start.data(), h_rs, (size_t)nreg * 4);
        memcpy(rg.end.data(), h_re, (size_t)nreg * 4);
        memcpy(rg.a.data(), h_ra, (size_t)nreg * 4);
        memcpy(rg.b.data(), h_rb, (size_t)nreg * 4);
        const uint32_t n_ent = hc->c[C_NENT];
        const uint64_t seeds_bytes = q_seedoff[nreg];
        DBuf<uint8_t> d_seeds, d_rech_pool;
        DBuf<uint32_t> d_ent_order, d_ent_len;
        DBuf<uint64_t> d_ent_poff;
        d_seeds.alloc(seeds_bytes + 1, s);
        d_rech_pool.alloc(rech_bytes + 1, s);
        d_ent_order.alloc(std::max(n_ent, 1u), s);
        d_ent_len.alloc(std::max(n_ent, 1u), s);
        d_ent_poff.alloc(std::max(n_ent, 1u), s);
        // The re-check reads the DP bases only next to RECH regions (k-1 flank bases, the stretch between chained
        // regions): one window [a - W, b + W) per RECH region is gathered on the device instead of downloading all N.
        const uint32_t kWin = 128;
        std::vector<uint32_t> win_lo, win_r;
        std::vector<uint64_t> win_off(1, 0);
        for (uint32_t r = nreg; r-- > 0;)  // ascending position
            if (lab[r] & LABLE_RECH) {
                const uint32_t lo = rg.a[r] > kWin ? rg.a[r] - kWin : 0, hi = (uint32_t)std::min<uint64_t>((uint64_t)rg.b[r] + kWin, N_true);
                win_lo.push_back(lo);
                win_r.push_back(r);
                win_off.push_back(win_off.back() + (hi - lo));
            }
        const uint32_t n_win = (uint32_t)win_lo.size();
        bool full_cbase = win_off.back() > N_true / 2;
        DBuf<uint32_t> d_win_lo;
        DBuf<uint64_t> d_win_off;
        DBuf<uint8_t> d_win;
        h = timer.begin("seed_gather", 3);
        if (full_cbase) {
            d_cbase.download(p_cbase.p, N_true);
        } else if (n_win) {
            d_win_lo.alloc(n_win, s);
            d_win_off.alloc(n_win + 1, s);
            d_win.alloc(win_off.back() + 1, s);
            Uploader{sc->p_h2d, sc->h2d_used, s}.put(d_win_lo.p, win_lo.data(), n_win);
            Uploader{sc->p_h2d, sc->h2d_used, s}.put(d_win_off.p, win_off.data(), (size_t)n_win + 1);
            gather_ranges(d_cbase.p, d_win_lo.p, d_win_off.p, nullptr, n_win, d_win.p, s);
            h_win.resize(win_off.back() + 16);
            d_win.download(h_win.data(), win_off.back());
            h2d += (uint64_t)n_win * 12;
        }
        assemble_seed_gather(ad, d_seeds.p, nreg, s);
        if (n_ent) rech_gather(g, d_ent_off.p, d_rech_boff.p, d_ent_order.p, d_ent_len.p, d_ent_poff.p, d_rech_pool.p, nreg, s);
        timer.end(h);
        // round 2: the seed strings and the survivors of the RECH regions
        h_seeds.resize(seeds_bytes + 16);
        h_rech_pool.resize(rech_bytes + 16);
        d_seeds.download(h_seeds.data(), seeds_bytes);
        if (rech_bytes) d_rech_pool.download(h_rech_pool.data(), rech_bytes);
        std::vector<uint32_t> ent_order(n_ent), ent_len(n_ent);
        std::vector<uint64_t> ent_poff(n_ent);
        if (n_ent) {
            d_ent_order.download(ent_order.data(), n_ent);
            d_ent_len.download(ent_len.data(), n_ent);
            d_ent_poff.download(ent_poff.data(), n_ent);
        }
        NP2_CUDA(cudaStreamSynchronize(s));
        n_sync++;
        timer.hend("host:seed_sync2");
        d2h += (full_cbase ? (uint64_t)N_true : win_off.back()) + seeds_bytes + rech_bytes + (uint64_t)nreg * 49 + (uint64_t)n_ent * 16;
        if (!full_cbase && n_win) {
            for (uint32_t w = 0; w < n_win; w++)
                memcpy(p_cbase.p + win_lo[w], h_win.data() + win_off[w], win_off[w + 1] - win_off[w]);
            // Exactness check: every DP base the re-check can touch must lie inside a window.  Left/right flanks take at
            // most kmax-1 DP bases walking over neighbouring regions (their alleles only shorten the walk); chained
            // regions (closer than kmax in position) read the whole stretch between them.
            const uint32_t need = tables.back()->dev.k - 1;
            bool ok = true;
            for (uint32_t w = 0; w < n_win && ok; w++) {
                const uint32_t r = win_r[w];
                const uint64_t wlo = win_lo[w], whi = wlo + (win_off[w + 1] - win_off[w]);
                uint32_t got = 0, rr = r;  // left walk (r grows towards lower positions)
                uint64_t i = rg.a[r];
                for (;;) {
                    const uint64_t lo = rr + 1 < nreg ? rg.b[rr + 1] : 0;
                    const uint64_t take = std::min<uint64_t>(need - got, i - lo);
                    if (i - take < wlo) ok = false;
                    got += (uint32_t)take;
                    if (got >= need || rr + 1 >= nreg) break;
                    rr++;
                    i = rg.a[rr];
                }
                got = 0, rr = r, i = rg.b[r];
                for (;;) {
                    const uint64_t hi = rr > 0 ? rg.a[rr - 1] : N_true;
                    const uint64_t take = std::min<uint64_t>(need - got, hi - i);
                    if (i + take > whi) ok = false;
                    got += (uint32_t)take;
                    if (got >= need || rr == 0) break;
                    rr--;
                    i = rg.b[rr];
                }
                if (w + 1 < n_win) {
                    const uint32_t r2 = win_r[w + 1];
                    if (rg.start[r2] < rg.end[r] + need + 1 && rg.a[r2] > whi && win_lo[w + 1] > whi) ok = false;
                }
            }
            if (!ok) {  // pathological layout (very long insertions next to a RECH region): take everything
                d_cbase.download(p_cbase.p, N_true);
                NP2_CUDA(cudaStreamSynchronize(s));
                n_sync++;
                d2h += N_true;
            }
        }
        // patched view, regions in ascending position (q = nreg - 1 - r)
        pc.reset(nreg);
        pc.cbase = p_cbase.p;
        pc.N = N_true;
        {
            uint64_t rb = 0;
            for (uint32_t r = 0; r < nreg; r++) {
                const uint32_t q = nreg - 1 - r;
                pc.start[q] = rg.start[r];
                pc.end[q] = rg.end[r];
                pc.a[q] = rg.a[r];
                pc.b[q] = rg.b[r];
                pc.lable[q] = lab[r];
                pc.seed[q].s = h_seeds.data() + q_seedoff[q];
                pc.seed[q].len = seed_len[r];
                pc.seed[q].dev_off = seed_off[r];
                for (uint32_t x = 0; x < nsurv[r]; x++) {  // rech pool is laid out in r order, survivors in rank order
                    Allele al;
                    al.s = h_rech_pool.data() + rb;
                    al.len = ent_len[ent_off[r] + x];
                    al.order = ent_order[ent_off[r] + x];
                    al.dev_off = ent_poff[ent_off[r] + x];
                    rb += al.len;
                    pc.cand[q].push_back(al);
                }
            }
        }
    }
    timer.hend("host:seed_view");
    bool changed = false;
    for (size_t ti = 0; ti < tables.size(); ti++) {
        Reupdate ru;
        reupdate_build(pc, tables[ti]->dev.k, ru);
        timer.hend("host:reupdate_build");
        const size_t ns = ru.off.size() - 1;
        std::vector<uint16_t> ks(ns);
        if (ns) {
            DBuf<uint8_t> d_rp;
            DBuf<uint64_t> d_ro;
            DBuf<uint16_t> d_rk;
            d_rp.alloc(ru.pool.size() + 1, s);
            d_ro.alloc(ns + 1, s);
            d_rk.alloc(ns, s);
            Uploader{sc->p_h2d, sc->h2d_used, s}.put(d_rp.p, ru.pool.data(), ru.pool.size());
            Uploader{sc->p_h2d, sc->h2d_used, s}.put(d_ro.p, ru.off.data(), ns + 1);
            h = timer.begin("yak_seq_kscore", 1);
            seq_kscore(tables[ti]->dev, d_rp.p, d_ro.p, nullptr, ns, opt.min_kmer_count, d_rk.p, s);
            timer.end(h);
            d_rk.download(ks.data(), ns);
            NP2_CUDA(cudaStreamSynchronize(s));
            n_sync++;
            h2d += ru.pool.size() + (ns + 1) * 8;
            d2h += ns * 2;
            n_probes += ru.pool.size();
            changed = true;
        }
        timer.hend("host:reupdate_score_sync");
        reupdate_apply(pc, ru, ks.data(), (uint32_t)ti + 1);
        timer.hend("host:reupdate_apply");
    }
    if (dump) {
        dm_reg_lable.resize(nreg);
        for (uint32_t r = 0; r < nreg; r++) dm_reg_lable[r] = pc.lable[nreg - 1 - r];
    }
    /* ---- final consensus assembled on the device from the (possibly re-chosen) seeds */
    long long total_shift = 0;
    DBuf<uint32_t> d_ch_r, d_ch_len;
    DBuf<uint64_t> d_ch_off;
    if (sparse_view) {  // only the seeds the re-check changed go back to the device
        std::vector<uint32_t> ch_r, ch_len;
        std::vector<uint64_t> ch_off;
        total_shift = sv_shift0;
        const uint32_t nsub = (uint32_t)sv_sub.size();
        for (uint32_t q = 0; q < nsub; q++) {
            const Allele &al = pc.seed[q];
            if (al.dev_off == sv_init_off[q] && al.len == sv_init_len[q]) continue;
            ch_r.push_back(sv_sub[nsub - 1 - q]);
            ch_off.push_back(al.dev_off);
            ch_len.push_back(al.len);
            total_shift += (long long)al.len - (long long)sv_init_len[q];
        }
        if (!ch_r.empty()) {
            const uint32_t nc = (uint32_t)ch_r.size();
            d_ch_r.alloc(nc, s);
            d_ch_len.alloc(nc, s);
            d_ch_off.alloc(nc, s);
            Uploader upl{sc->p_h2d, sc->h2d_used, s};
            upl.put(d_ch_r.p, ch_r.data(), nc);
            upl.put(d_ch_len.p, ch_len.data(), nc);
            upl.put(d_ch_off.p, ch_off.data(), nc);
            seed_scatter(nc, d_ch_r.p, d_ch_off.p, d_ch_len.p, d_r_seed_off.p, d_r_seed_len.p, s);
            h2d += (uint64_t)nc * 16;
        }
    } else {
        if (changed) {
            for (uint32_t r = 0; r < nreg; r++) {
                const Allele &al = pc.seed[nreg - 1 - r];
                seed_off[r] = al.dev_off;
                seed_len[r] = al.len;
            }
            d_r_seed_off.upload(seed_off, nreg);
            d_r_seed_len.upload(seed_len, nreg);
            h2d += (uint64_t)nreg * 12;
        }
        for (uint32_t q = 0; q < nreg; q++) total_shift += (long long)pc.seed[q].len - (long long)(pc.b[q] - pc.a[q]);
    }
    const uint64_t out_n = (uint64_t)((long long)N_true + total_shift);
    DBuf<uint8_t> d_out;
    d_out.alloc(out_n + 1, s);
    res_base.resize(std::max<uint64_t>(out_n, 1));
    res_base.n = out_n;
    h = timer.begin("assemble", 3);
    assemble_sizes(ad, nreg, s);
    region_scan_i64(d_q_delta.p, d_q_shift.p, nreg, -1, cd, sp, s);
    assemble_final(ad, d_out.p, nreg, s);
    timer.end(h);
    timer.hend("host:assemble_prep");
    d_out.download(res_base.p, out_n);
    NP2_CUDA(cudaStreamSynchronize(s));
    n_sync++;
    d2h += out_n;
    // FASTA header span (main.rs:627-632)
    const size_t nview = pc.a.size();  // the sparse view always holds the first and the last region
    res_first = (pc.a[0] == 0) ? pc.start[0] : edge_pos[0];
    res_last = (pc.b[nview - 1] == N_true) ? pc.start[nview - 1] : edge_pos[1];
    res_pos_valid = false;
    res_sparse = sparse_view;
    if (sparse_view) {  // positions are produced on request from the device copies (np2_job_get_consensus)
        res_nreg = nreg;
        jd_reg_start.alloc(nreg, s);
        jd_reg_a.alloc(nreg, s);
        jd_reg_b.alloc(nreg, s);
        jd_reg_len.alloc(nreg, s);
        NP2_CUDA(cudaMemcpyAsync(jd_reg_start.p, d_rstart.p, (size_t)nreg * 4, cudaMemcpyDeviceToDevice, s));
        NP2_CUDA(cudaMemcpyAsync(jd_reg_a.p, d_ra.p, (size_t)nreg * 4, cudaMemcpyDeviceToDevice, s));
        NP2_CUDA(cudaMemcpyAsync(jd_reg_b.p, d_rb.p, (size_t)nreg * 4, cudaMemcpyDeviceToDevice, s));
        NP2_CUDA(cudaMemcpyAsync(jd_reg_len.p, d_r_seed_len.p, (size_t)nreg * 4, cudaMemcpyDeviceToDevice, s));
    }
    timer.hend("host:assemble");
    return iter + 1;
  }
}

void np2_job::run(int32_t dump_it) {
    dump_iter = dump_it;
    cudaStream_t s = ctx->stream;
    res_base.n = 0;
    res_pos.clear();
    res_pos_valid = false;
    res_patch.reset(0);
    res_sparse = false;
    timer.s = s;
    timer.reset();
    const unsigned long long launches0 = launch_counter();
    n_launch = 0;
    n_probes = 0;
    n_sync = 0;
    memset(stats, 0, sizeof stats);
    dm_dropped.clear();
    dm_pair_key.clear();
    dm_pair_val.clear();
    dm_msa_off.clear();
    dm_msa_bases.clear();
    dm_msa_delta.clear();
    dm_msa_count.clear();
    dm_msa_besti.clear();
    dm_can_roff.clear();
    dm_can_order.clear();
    dm_can_kscore.clear();
    dm_can_kmer.clear();
    dm_can_seq_off.clear();
    dm_can_seq.clear();
    dm_reg_start.clear();
    dm_reg_end.clear();
    dm_reg_lable.clear();
    if (L < opt.min_ctg_len) {  // main.rs:1727-1730
        res_base.resize(std::max(L, 1u));
        res_base.n = L;
        memcpy(res_base.p, tseq.data(), L);
        res_pos.resize(L);
        for (uint32_t p = 0; p < L; p++) res_pos[p] = p;
        res_pos_valid = true;
        res_first = 0;
        res_last = L ? L - 1 : 0;
        return;
    }
    if (!uploaded) upload();
    const uint32_t n = R.n_reads;
    sc->arena.reset(s);
    ArenaScope arena_scope(&sc->arena);
    if (!sc->d_counts) NP2_CUDA(cudaMalloc((void **)&sc->d_counts, sizeof(CountsHost)));
    sc->p_counts.resize(sizeof(CountsHost) + 16);
    hc = reinterpret_cast<CountsHost *>(sc->p_counts.p);
    cd.c = sc->d_counts;
    cd.q = reinterpret_cast<unsigned long long *>(sc->d_counts + C_COUNT);
    // tickets + tile descriptors of one pass: ~10 position-sized scans, a few record- and region-sized ones
    sc->scan_pool.reserve(12 * ((size_t)L / kScanTile + 2) + 4 * ((size_t)ing.total_cols / 8 / kScanTile + 2) + 8192, s);
    sc->scan_pool.begin(s);
    // page-locked buffers at their final size right away (see PBuf::reserve)
    sc->p_h2d.reserve((size_t)n * 24 + (4u << 20));
    sc->h2d_used = 0;
    p_cbase.reserve((size_t)L + L / 4 + 65536);
    res_base.reserve((size_t)L + L / 4 + 65536);
    Uploader up{sc->p_h2d, sc->h2d_used, s};
    timer.trace_begin();
    const int h_total = timer.begin("total", 0);
    int h = timer.begin("trim_scan", 1);
    trim_scan(R, d_ref.p, L, s);
    timer.end(h);
    // The per-read results of the trim go to the host on the copy stream while the main stream packs the columns; the
    // host decides which reads are kept (ingest_finish) during that kernel.
    if (!sc->ev_trim) NP2_CUDA(cudaEventCreateWithFlags(&sc->ev_trim, cudaEventDisableTiming));
    NP2_CUDA(cudaEventRecord(sc->ev_trim, s));
    h = timer.begin("pack_columns", 1);  // a single kernel; the plain blocks' "not all reference" bits fall out of it
    d_blk_odd.alloc(((size_t)ing.ck_off.back() + 255) / 256 * 8 + 8, s);
    const bool flags_done = pack_columns(R, d_ref.p, ing.ck_off.back(), d_refpk.p, d_blk_odd.p, s);
    timer.end(h);
    // The per-stripe block lists only need the trim and the packed columns: enqueued here, they run while the host waits
    // for the trim results and decides which reads are kept.
    max_span = 0;
    for (uint32_t v : ing.rspan) max_span = std::max(max_span, v);
    const uint32_t n_stripes = pileup_stripes(L);
    d_first_ge.alloc(n_stripes + 1, s);
    stripe_reads(R, L, d_first_ge.p, s);
    h = timer.begin("block_flags", 4);
    if (!flags_done) block_flags(R, ing.ck_off.back(), d_code.p, d_refpk.p, d_blk_odd.p, s);
    // per stripe: which of its blocks K2 has to walk (a property of the reads and the contig, like the flags)
    DBuf<uint32_t> d_odd_cnt;
    d_odd_cnt.alloc(n_stripes, s);
    d_odd_off.alloc(n_stripes + 1, s);
    stripe_odd_count(R, d_blk_odd.p, d_first_ge.p, L, max_span, d_odd_cnt.p, s);
    stripe_odd_offsets(d_odd_cnt.p, d_odd_off.p, L, nullptr, sc->scan_pool, s);
    timer.end(h);
    NP2_CUDA(cudaMemcpyAsync(&hc->c[0], d_odd_off.p + n_stripes, 4, cudaMemcpyDeviceToHost, s));
    {
        cudaStream_t c2 = ctx->copy_stream;
        NP2_CUDA(cudaStreamWaitEvent(c2, sc->ev_trim, 0));
        sc->p_trim.resize((size_t)std::max(n, 1u) * 12 + 16);
        uint32_t *pt = reinterpret_cast<uint32_t *>(sc->p_trim.p);
        if (n) {
            NP2_CUDA(cudaMemcpyAsync(pt, d_ts.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c2));
            NP2_CUDA(cudaMemcpyAsync(pt + n, d_te.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c2));
            NP2_CUDA(cudaMemcpyAsync(pt + 2 * (size_t)n, d_n.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c2));
        }
        NP2_CUDA(cudaMemcpyAsync(pt + 3 * (size_t)n, d_bad.p, 4, cudaMemcpyDeviceToHost, c2));
        timer.hbegin();
        NP2_CUDA(cudaStreamSynchronize(c2));
        timer.hend("host:trim_sync");
        n_sync++;
        h_ts.assign(pt, pt + n);
        h_te.assign(pt + n, pt + 2 * (size_t)n);
        h_n.assign(pt + 2 * (size_t)n, pt + 3 * (size_t)n);
        d2h += (uint64_t)n * 12;
        if (pt[3 * (size_t)n]) throw np2::Error(NP2_ERR_FORMAT, "contig holds a byte the reference cannot index (>= 128 or '-')");
    }
    ingest_finish();
    timer.hend("host:ingest_finish");
    up.put(d_blank.p, h_blank.data(), std::max(n, 1u));
    d_order.alloc(std::max(n, 1u), s);
    if (n) up.put(d_order.p, read_order.data(), n);
    DBuf<uint32_t> d_as_te, d_W;
    const uint32_t na = (uint32_t)as_read.size();
    if (opt.iter_count > 1) {  // slot ranges of the pair accumulator: windows + exclusive scan, all on the device
        d_as_pos.alloc(na, s);
        d_as_te.alloc(na, s);
        d_W.alloc(na + 1, s);
        d_pair_off.alloc(na + 1, s);
        up.put(d_as_pos.p, h_as_pos.data(), na);
        up.put(d_as_te.p, h_as_te.data(), na);
        geno_pair_windows(d_as_pos.p, d_as_te.p, na, d_W.p, s);
        geno_pair_window_offsets(d_W.p, d_pair_off.p, na, sc->scan_pool, s);
        NP2_CUDA(cudaMemcpyAsync(&hc->q[0], d_pair_off.p + na, 8, cudaMemcpyDeviceToHost, s));
    }
    timer.hbegin();
    NP2_CUDA(cudaStreamSynchronize(s));
    timer.hend("host:stripe_list_sync");
    n_sync++;
    if (opt.iter_count > 1) pair_slots = hc->q[0];
    const uint32_t n_odd_list = hc->c[0];
    d_odd_list.alloc(std::max(n_odd_list, 1u), s);
    h = timer.begin("block_flags", 1);
    stripe_odd_fill(R, d_blk_odd.p, d_first_ge.p, L, max_span, d_odd_off.p, d_odd_list.p, n_odd_list, s);
    timer.end(h);

    if (dump_iter >= 0) {  // reads as the oracle reports them (after the clip filter)
        std::vector<uint8_t> nib(ing.nib_off.back() + 16);
        d_nib.download(nib.data(), nib.size());
        NP2_CUDA(cudaStreamSynchronize(s));
        d_rec_idx.clear();
        dm_ts.clear();
        dm_te.clear();
        dm_blank.clear();
        dm_nib.clear();
        dm_nib_off.assign(1, 0);
        for (size_t a = 0; a < as_read.size(); a++) {
            if (a == 0) {
                d_rec_idx.push_back(-1);
                dm_ts.push_back(0);
                dm_te.push_back(L - 1);
                dm_blank.push_back(0);
            } else {
                const uint32_t i = (uint32_t)as_read[a];
                d_rec_idx.push_back(ing.rec_idx[i]);
                dm_ts.push_back(h_ts[i]);
                dm_te.push_back(h_te[i]);
                dm_blank.push_back(h_blank[i]);
                if (!h_blank[i]) {
                    const uint32_t nn = h_n[i];
                    const uint8_t *src = nib.data() + ing.nib_off[i];
                    const size_t bytes = ((size_t)nn + 1) / 2 + 1;
                    size_t o = dm_nib.size();
                    dm_nib.insert(dm_nib.end(), src, src + bytes);
                    if (nn & 1) dm_nib[o + bytes - 1] = 0;  // Vec is zero-initialised past the terminator nibble
                }
            }
            dm_nib_off.push_back(dm_nib.size());
        }
    }
    for (uint32_t it = 0; it < opt.iter_count;) it = iteration(it);
    timer.end(h_total);
    n_launch = launch_counter() - launches0;
    NP2_CUDA(cudaStreamSynchronize(s));
    timer.mark("end");
    g_trace = nullptr;
    timer.trace_dump(this);
    timer.collect();
    hc = nullptr;
}

/* ================================================================= C ABI */

extern "C" {

const char *np2_last_error(void) { return g_err.c_str(); }

void np2_opts_default(np2_opts *o) {
    memset(o, 0, sizeof *o);
    o->min_kmer_count = 5;
    o->iter_count = 2;
    o->model = 0;
    o->min_read_len = 1000;
    o->min_ctg_len = 1000000;
    o->max_indel_len = 20;
    o->min_map_len = 500;
    o->min_map_fra = 0.5f;
    o->min_map_qual = 1;
    o->max_clip_len = 100;
}

int np2_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int np2_ctx_create(int device, np2_ctx **out) {
    return guard([&] {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n == 0)
            throw np2::Error(NP2_ERR_CUDA, std::string("no CUDA device: libnp2gpu has no CPU fallback (") +
                                               cudaGetErrorString(e) + ")");
        if (device < 0 || device >= n) throw np2::Error(NP2_ERR_ARG, "bad device index");
        NP2_CUDA(cudaSetDevice(device));
        np2::host_alloc_hook = [](size_t n) -> void * {
            void *p = nullptr;
            return cudaHostAlloc(&p, n, cudaHostAllocPortable) == cudaSuccess ? p : nullptr;
        };
        np2::host_free_hook = [](void *p) { cudaFreeHost(p); };
        np2_ctx *c = new np2_ctx();
        c->device = device;
        NP2_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        {
            int lo = 0, hi = 0;  // numerically lower = higher priority
            NP2_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            NP2_CUDA(cudaStreamCreateWithPriority(&c->copy_stream, cudaStreamNonBlocking, hi));
        }
        // a private pool that keeps freed blocks: per-iteration scratch is re-used instead of going back to the driver
        cudaMemPoolProps props;
        memset(&props, 0, sizeof props);
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        NP2_CUDA(cudaMemPoolCreate(&c->pool, &props));
        uint64_t thr = UINT64_MAX;
        NP2_CUDA(cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &thr));
        {
            std::lock_guard<std::mutex> lk(g_pool_mu);
            g_stream_pool[c->stream] = c->pool;
        }
        *out = c;
    });
}
void np2_ctx_destroy(np2_ctx *ctx) {
    if (ctx) ctx_release(ctx);
}

int np2_yak_load(np2_ctx *ctx, const char *path, np2_table **out) {
    return guard([&] {
        NP2_CUDA(cudaSetDevice(ctx->device));
        FILE *fp = fopen(path, "rb");
        if (!fp) throw np2::Error(NP2_ERR_IO, std::string("cannot open ") + path);
        std::unique_ptr<FILE, int (*)(FILE *)> fc(fp, fclose);
        uint8_t hdr[16];
        if (fread(hdr, 1, 16, fp) != 16 || memcmp(hdr, "YAK\2", 4) != 0)
            throw np2::Error(NP2_ERR_FORMAT, "The input binary k-mer dump file is incompatible.");  // kmer.rs:76-80
        uint32_t k, pre, cb;
        memcpy(&k, hdr + 4, 4);
        memcpy(&pre, hdr + 8, 4);
        memcpy(&cb, hdr + 12, 4);
        if (cb != 10) throw np2::Error(NP2_ERR_FORMAT, "different YAK_COUNTER_BITS");  // kmer.rs:90
        if (pre != 10) throw np2::Error(NP2_ERR_UNSUPPORTED, "yak prefix bits must be 10 (kmer.rs:52-54 vs htab.c:59)");
        if (k == 0 || k > 63) throw np2::Error(NP2_ERR_FORMAT, "bad k in yak header");
        fseek(fp, 0, SEEK_END);
        const uint64_t fsz = (uint64_t)ftell(fp);
        fseek(fp, 16, SEEK_SET);
        std::vector<uint64_t> keys;
        keys.reserve(fsz / 8);
        std::vector<uint32_t> sub_off(1025, 0);
        uint64_t max_sub = 0;
        for (uint32_t b = 0; b < 1024; b++) {
            uint32_t cs[2];
            if (fread(cs, 4, 2, fp) != 2) throw np2::Error(NP2_ERR_FORMAT, "Failed to parse the dump file");
            size_t o = keys.size();
            keys.resize(o + cs[1]);
            if (cs[1] && fread(keys.data() + o, 8, cs[1], fp) != cs[1])
                throw np2::Error(NP2_ERR_FORMAT, "Failed to parse the dump file");
            if (keys.size() >= (1ull << 32)) throw np2::Error(NP2_ERR_UNSUPPORTED, "more than 2^32 keys: use np2_yak_from_arrays in chunks");
            sub_off[b + 1] = (uint32_t)keys.size();
            max_sub = std::max<uint64_t>(max_sub, cs[1]);
        }
        std::unique_ptr<np2_table> t(new np2_table());
        table_alloc(ctx, t.get(), k, keys.size(), max_sub);
        DBuf<uint64_t> d_keys;
        DBuf<uint32_t> d_off;
        DBuf<int> d_err;
        d_keys.alloc(std::max<size_t>(keys.size(), 1), ctx->stream);
        d_off.alloc(1025, ctx->stream);
        d_err.alloc(1, ctx->stream);
        d_err.zero();
        if (!keys.empty()) d_keys.upload(keys.data(), keys.size());
        d_off.upload(sub_off.data(), 1025);
        table_insert_filekeys(t->dev, d_keys.p, d_off.p, keys.size(), d_err.p, ctx->stream);
        int err = 0;
        d_err.download(&err, 1);
        NP2_CUDA(cudaStreamSynchronize(ctx->stream));
        if (err) throw np2::Error(NP2_ERR_INTERNAL, "table insertion overflow");
        *out = t.release();
    });
}

int np2_yak_from_arrays(np2_ctx *ctx, uint32_t k, const uint64_t *hashes, const uint16_t *counts, uint64_t n,
                        np2_table **out) {
    return guard([&] {
        NP2_CUDA(cudaSetDevice(ctx->device));
        if (k == 0 || k > 63) throw np2::Error(NP2_ERR_ARG, "bad k");
        std::vector<uint64_t> cnt(1024, 0);
        for (uint64_t i = 0; i < n; i++) cnt[hashes[i] & 1023]++;
        uint64_t max_sub = *std::max_element(cnt.begin(), cnt.end());
        std::unique_ptr<np2_table> t(new np2_table());
        table_alloc(ctx, t.get(), k, n, max_sub);
        DBuf<uint64_t> d_h;
        DBuf<uint16_t> d_c;
        DBuf<int> d_err;
        d_h.alloc(std::max<uint64_t>(n, 1), ctx->stream);
        d_c.alloc(std::max<uint64_t>(n, 1), ctx->stream);
        d_err.alloc(1, ctx->stream);
        d_err.zero();
        if (n) {
            d_h.upload(hashes, n);
            d_c.upload(counts, n);
        }
        table_insert(t->dev, d_h.p, d_c.p, n, d_err.p, ctx->stream);
        int err = 0;
        d_err.download(&err, 1);
        NP2_CUDA(cudaStreamSynchronize(ctx->stream));
        if (err) throw np2::Error(NP2_ERR_INTERNAL, "table insertion overflow");
        *out = t.release();
    });
}

/* ---- table replicas for the other GPUs of the box */
static np2_table *table_shell(np2_ctx *ctx, uint32_t k, uint64_t n, uint32_t nb) {
    std::unique_ptr<np2_table> t(new np2_table());
    t->ctx = ctx;
    t->dev.k = k;
    t->dev.n = n;
    t->dev.nb = nb;
    t->bytes = 1024ull * nb * kBucketSlots * 8;
    NP2_CUDA(cudaMalloc((void **)&t->dev.slots, t->bytes));
    ctx->refs++;
    return t.release();
}
int np2_yak_clone(np2_ctx *dst_ctx, const np2_table *src, np2_table **out) {
    return guard([&] {
        if (!dst_ctx || !src || !out) throw np2::Error(NP2_ERR_ARG, "null argument");
        NP2_CUDA(cudaSetDevice(src->ctx->device));
        NP2_CUDA(cudaStreamSynchronize(src->ctx->stream));  // the source image is complete
        NP2_CUDA(cudaSetDevice(dst_ctx->device));
        np2_table *t = table_shell(dst_ctx, src->dev.k, src->dev.n, src->dev.nb);
        cudaError_t e = cudaMemcpyPeerAsync(t->dev.slots, dst_ctx->device, src->dev.slots, src->ctx->device, t->bytes,
                                            dst_ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(dst_ctx->stream);
        if (e != cudaSuccess) {
            np2_yak_free(t);
            throw np2::Error(NP2_ERR_CUDA, std::string("peer copy of the table image: ") + cudaGetErrorString(e));
        }
        *out = t;
    });
}
int np2_yak_image(const np2_table *t, const void **d_image, uint64_t *bytes, uint32_t *buckets_per_subtable) {
    return guard([&] {
        if (!t || !d_image || !bytes || !buckets_per_subtable) throw np2::Error(NP2_ERR_ARG, "null argument");
        NP2_CUDA(cudaSetDevice(t->ctx->device));
        NP2_CUDA(cudaStreamSynchronize(t->ctx->stream));
        *d_image = t->dev.slots;
        *bytes = t->bytes;
        *buckets_per_subtable = t->dev.nb;
    });
}
int np2_yak_adopt(np2_ctx *ctx, uint32_t k, uint64_t n_keys, uint32_t buckets_per_subtable, const void *d_image,
                  uint64_t bytes, np2_table **out) {
    return guard([&] {
        if (!ctx || !d_image || !out) throw np2::Error(NP2_ERR_ARG, "null argument");
        if (k == 0 || k > 63 || buckets_per_subtable == 0) throw np2::Error(NP2_ERR_ARG, "bad table geometry");
        if (bytes != 1024ull * buckets_per_subtable * kBucketSlots * 8)
            throw np2::Error(NP2_ERR_ARG, "image size does not match buckets_per_subtable");
        NP2_CUDA(cudaSetDevice(ctx->device));
        np2_table *t = table_shell(ctx, k, n_keys, buckets_per_subtable);
        cudaError_t e = cudaMemcpyAsync(t->dev.slots, d_image, bytes, cudaMemcpyDeviceToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            np2_yak_free(t);
            throw np2::Error(NP2_ERR_CUDA, std::string("copy of the table image: ") + cudaGetErrorString(e));
        }
        *out = t;
    });
}

/* ---- yak count on the device (np2_count.cu) */
struct np2_counter {
    np2_ctx *ctx = nullptr;
    uint32_t k = 0;
    np2::KmerCounts acc;
    uint64_t n_kmers = 0;
};
int np2_count_create(np2_ctx *ctx, uint32_t k, np2_counter **out) {
    return guard([&] {
        if (!ctx || !out) throw np2::Error(NP2_ERR_ARG, "null argument");
        if (k == 0 || k > 63) throw np2::Error(NP2_ERR_ARG, "-k must be smaller than 64");  // yak/main.c:57-60
        np2_counter *c = new np2_counter();
        c->ctx = ctx;
        c->k = k;
        ctx->refs++;
        *out = c;
    });
}
void np2_count_destroy(np2_counter *c) {
    if (!c) return;
    cudaSetDevice(c->ctx->device);
    np2::count_free(c->acc, c->ctx->stream);
    cudaStreamSynchronize(c->ctx->stream);
    np2_ctx *ctx = c->ctx;
    delete c;
    ctx_release(ctx);
}
int np2_count_add(np2_counter *c, const uint8_t *seqs, const uint64_t *seq_off, uint64_t n_seqs) {
    return guard([&] {
        if (!c || (!seqs && n_seqs) || !seq_off) throw np2::Error(NP2_ERR_ARG, "null argument");
        NP2_CUDA(cudaSetDevice(c->ctx->device));
        cudaStream_t s = c->ctx->stream;
        // batches of <= 1 Gbase: reads joined by 'N' (any non-ACGT byte restarts yak's k-mer window, count.c:41)
        const uint64_t kBatch = 1ull << 30;
        PBuf<uint8_t> stage;
        DBuf<uint8_t> d_seq;
        uint64_t i = 0;
        while (i < n_seqs) {
            uint64_t j = i, bytes = 0;
            while (j < n_seqs && (j == i || bytes + (seq_off[j + 1] - seq_off[j]) + 1 <= kBatch)) {
                bytes += seq_off[j + 1] - seq_off[j] + 1;
                j++;
            }
            if (bytes >= (1ull << 31)) throw np2::Error(NP2_ERR_UNSUPPORTED, "a single sequence of 2^31 bases or more");
            stage.resize(bytes);
            uint64_t w = 0;
            for (uint64_t r = i; r < j; r++) {
                const uint64_t l = seq_off[r + 1] - seq_off[r];
                memcpy(stage.p + w, seqs + seq_off[r], l);
                stage.p[w + l] = 'N';
                w += l + 1;
            }
            d_seq.alloc(bytes, s);
            d_seq.upload(stage.p, bytes);
            np2::count_add(c->acc, d_seq.p, bytes, c->k, &c->n_kmers, s);
            i = j;
        }
    });
}
uint64_t np2_count_distinct(const np2_counter *c, uint64_t *n_kmers) {
    if (n_kmers) *n_kmers = c->n_kmers;
    return c->acc.n;
}
int np2_count_finish(np2_counter *c, uint32_t min_count, const char *dump_path, np2_table **out_table) {
    return guard([&] {
        if (!c) throw np2::Error(NP2_ERR_ARG, "null argument");
        NP2_CUDA(cudaSetDevice(c->ctx->device));
        cudaStream_t s = c->ctx->stream;
        uint64_t *d_keys = nullptr, n = 0;
        uint16_t *d_cnt = nullptr;
        uint32_t sub[1024];
        np2::count_filter(c->acc, std::max(min_count, 1u), &d_keys, &d_cnt, &n, sub, s);
        struct Free {
            uint64_t *k;
            uint16_t *c;
            cudaStream_t s;
            ~Free() {
                cudaFreeAsync(k, s);
                cudaFreeAsync(c, s);
            }
        } fr{d_keys, d_cnt, s};
        if (dump_path) {  // yak_ch_dump (yak/htab.c:190-211)
            std::vector<uint64_t> fk(n);
            np2::count_file_keys(d_keys, d_cnt, n, fk.data(), s);
            FILE *fp = fopen(dump_path, "wb");
            if (!fp) throw np2::Error(NP2_ERR_IO, std::string("cannot write ") + dump_path);
            std::unique_ptr<FILE, int (*)(FILE *)> fc(fp, fclose);
            const uint32_t hdr[3] = {c->k, 10, 10};
            bool ok = fwrite("YAK\2", 1, 4, fp) == 4 && fwrite(hdr, 4, 3, fp) == 3;
            uint64_t o = 0;
            for (int b = 0; b < 1024 && ok; b++) {
                uint32_t cap = 4;  // khashl keeps its load below 0.75
                while ((uint64_t)cap * 3 / 4 < sub[b]) cap <<= 1;
                const uint32_t t[2] = {cap, sub[b]};
                ok = fwrite(t, 4, 2, fp) == 2 && fwrite(fk.data() + o, 8, sub[b], fp) == sub[b];
                o += sub[b];
            }
            if (!ok) throw np2::Error(NP2_ERR_IO, std::string("short write to ") + dump_path);
        }
        if (out_table) {
            std::unique_ptr<np2_table> t(new np2_table());
            table_alloc(c->ctx, t.get(), c->k, n, *std::max_element(sub, sub + 1024));
            DBuf<int> d_err;
            d_err.alloc(1, s);
            d_err.zero();
            table_insert(t->dev, d_keys, d_cnt, n, d_err.p, s);
            int err = 0;
            d_err.download(&err, 1);
            NP2_CUDA(cudaStreamSynchronize(s));
            if (err) throw np2::Error(NP2_ERR_INTERNAL, "table insertion overflow");
            *out_table = t.release();
        }
        NP2_CUDA(cudaStreamSynchronize(s));
    });
}

void np2_yak_free(np2_table *t) {
    if (!t) return;
    cudaSetDevice(t->ctx->device);
    cudaStreamSynchronize(t->ctx->stream);
    cudaFree(t->dev.slots);
    np2_ctx *c = t->ctx;
    delete t;
    ctx_release(c);
}
uint32_t np2_yak_k(const np2_table *t) { return t->dev.k; }
uint64_t np2_yak_size(const np2_table *t) { return t->dev.n; }
uint64_t np2_yak_device_bytes(const np2_table *t) { return t->bytes; }

int np2_yak_lookup(np2_ctx *ctx, const np2_table *t, const uint64_t *hashes, uint64_t n, uint32_t min_count,
                   uint16_t *counts) {
    return guard([&] {
        NP2_CUDA(cudaSetDevice(ctx->device));
        if (!n) return;
        DBuf<uint64_t> d_h;
        DBuf<uint16_t> d_c;
        d_h.alloc(n, ctx->stream);
        d_c.alloc(n, ctx->stream);
        d_h.upload(hashes, n);
        table_probe(t->dev, d_h.p, n, min_count, d_c.p, ctx->stream);
        d_c.download(counts, n);
        NP2_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}

int np2_yak_lookup_device(np2_ctx *ctx, const np2_table *t, const uint64_t *d_hashes, uint64_t n, uint32_t min_count,
                          uint16_t *d_counts, uint32_t repeat, float *ms) {
    return guard([&] {
        NP2_CUDA(cudaSetDevice(ctx->device));
        if (repeat == 0) repeat = 1;
        cudaEvent_t a, b;
        NP2_CUDA(cudaEventCreate(&a));
        NP2_CUDA(cudaEventCreate(&b));
        NP2_CUDA(cudaEventRecord(a, ctx->stream));
        for (uint32_t r = 0; r < repeat; r++) table_probe(t->dev, d_hashes, n, min_count, d_counts, ctx->stream);
        NP2_CUDA(cudaEventRecord(b, ctx->stream));
        NP2_CUDA(cudaStreamSynchronize(ctx->stream));
        float t_ms = 0;
        cudaEventElapsedTime(&t_ms, a, b);
        cudaEventDestroy(a);
        cudaEventDestroy(b);
        NP2_CUDA(cudaGetLastError());
        if (ms) *ms = t_ms / repeat;
    });
}

int np2_bench_gather(np2_ctx *ctx, uint64_t buf_bytes, uint64_t n_loads, uint32_t block_bytes, uint32_t repeat, float *ms) {
    return guard([&] {
        NP2_CUDA(cudaSetDevice(ctx->device));
        if (block_bytes != 32 && block_bytes != 64 && block_bytes != 128) throw np2::Error(NP2_ERR_ARG, "block_bytes must be 32, 64 or 128");
        if (repeat == 0) repeat = 1;
        DBuf<uint64_t> buf, sink;
        const uint64_t n_sectors = buf_bytes / 128 * 4;
        buf.alloc(n_sectors * 4, ctx->stream);
        sink.alloc(1, ctx->stream);
        NP2_CUDA(cudaMemsetAsync(buf.p, 0x5A, n_sectors * 32, ctx->stream));
        gather32(buf.p, n_sectors, n_loads, 1, sink.p, ctx->stream, (int)block_bytes);  // warm-up
        cudaEvent_t a, b;
        NP2_CUDA(cudaEventCreate(&a));
        NP2_CUDA(cudaEventCreate(&b));
        NP2_CUDA(cudaEventRecord(a, ctx->stream));
        for (uint32_t r = 0; r < repeat; r++)
            gather32(buf.p, n_sectors, n_loads, 7 + r * n_loads, sink.p, ctx->stream, (int)block_bytes);
        NP2_CUDA(cudaEventRecord(b, ctx->stream));
        NP2_CUDA(cudaStreamSynchronize(ctx->stream));
        float t = 0;
        cudaEventElapsedTime(&t, a, b);
        cudaEventDestroy(a);
        cudaEventDestroy(b);
        NP2_CUDA(cudaGetLastError());
        *ms = t / repeat;
    });
}
int np2_bench_gather32(np2_ctx *ctx, uint64_t buf_bytes, uint64_t n_loads, uint32_t repeat, float *ms) {
    return np2_bench_gather(ctx, buf_bytes, n_loads, 32, repeat, ms);
}
int np2_l2_fetch_granularity(np2_ctx *ctx, uint32_t bytes, uint32_t *previous) {
    return guard([&] {
        NP2_CUDA(cudaSetDevice(ctx->device));
        size_t cur = 0;
        NP2_CUDA(cudaDeviceGetLimit(&cur, cudaLimitMaxL2FetchGranularity));
        if (previous) *previous = (uint32_t)cur;
        if (bytes) NP2_CUDA(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, bytes));
    });
}

int np2_seq_kscore(np2_ctx *ctx, const np2_table *t, const uint8_t *seqs, const uint64_t *seq_off, uint64_t n,
                   uint32_t min_count, uint16_t *kscore) {
    return guard([&] {
        NP2_CUDA(cudaSetDevice(ctx->device));
        if (!n) return;
        DBuf<uint8_t> d_s;
        DBuf<uint64_t> d_o;
        DBuf<uint16_t> d_k;
        d_s.alloc(seq_off[n] + 1, ctx->stream);
        d_o.alloc(n + 1, ctx->stream);
        d_k.alloc(n, ctx->stream);
        if (seq_off[n]) d_s.upload(seqs, seq_off[n]);
        d_o.upload(seq_off, n + 1);
        seq_kscore(t->dev, d_s.p, d_o.p, nullptr, n, min_count, d_k.p, ctx->stream);
        d_k.download(kscore, n);
        NP2_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}

static void bgzf_to_device(np2_ctx *ctx, const uint8_t *comp, uint64_t comp_len, const uint64_t *payload_off,
                           const uint32_t *payload_len, const uint32_t *isize, uint32_t n_members, DBuf<uint8_t> &d_out,
                           uint64_t &total, float *kernel_ms);

// The body of np2_job_create / np2_job_create_bgzf.  `records` fills j.ing and leaves in j.bam / j.bam_len the buffer the
// SEQ + CIGAR spans are gathered from (host memory, or the device's own copy of the inflated records); size_hint = the
// record bytes the pool is sized for (known before the records are, in the BGZF case from the ISIZE fields).
static int job_create_impl(np2_ctx *ctx, const uint8_t *tseq, uint32_t tlen, uint64_t size_hint,
                           const std::function<void(np2_job &)> &records, np2_table *const *tables, uint32_t n_tables,
                           const np2_opts *opts, np2_job **out) {
    return guard([&] {
        if (!ctx || !tseq || !opts || !out) throw np2::Error(NP2_ERR_ARG, "null argument");
        if (n_tables == 0) throw np2::Error(NP2_ERR_ARG, "Missing yak file!");
        if (opts->iter_count == 0) throw np2::Error(NP2_ERR_ARG, "iter_count must be >= 1");
        std::unique_ptr<np2_job> j(new np2_job(ctx));
        j->opt = *opts;
        for (uint32_t i = 0; i < n_tables; i++) {
            if (!tables[i] || tables[i]->ctx->device != ctx->device)
                throw np2::Error(NP2_ERR_ARG, "table is null or lives on another GPU than the context");
            j->tables.push_back(tables[i]);
        }
        std::stable_sort(j->tables.begin(), j->tables.end(),
                         [](np2_table *a, np2_table *b) { return a->dev.k < b->dev.k; });  // option.rs:238
        j->L = tlen;
        if (tlen >= opts->min_ctg_len) {
            if (tlen < 16) throw np2::Error(NP2_ERR_UNSUPPORTED, "contig shorter than 16 bp");
            if (tlen >= (1u << 30)) throw np2::Error(NP2_ERR_UNSUPPORTED, "contig >= 2^30 bp (main.rs:270)");
            NP2_CUDA(cudaSetDevice(ctx->device));
            {   // Grow the context's memory pool in ONE step to what a contig of this size needs (record bytes ~ 1.5 x
                // alignment columns; ~6 B per column + ~100 B per position of device state): a cold pool otherwise grows
                // by a hundred small mappings, which costs more than the whole polish.
                const uint64_t want = size_hint * 4 + (uint64_t)tlen * 100 + (64ull << 20);
                if (want > ctx->pool_warm) {
                    void *p = nullptr;
                    if (cudaMallocFromPoolAsync(&p, want, ctx->pool, ctx->stream) == cudaSuccess) {
                        cudaFreeAsync(p, ctx->stream);
                        ctx->pool_warm = want;
                    } else {
                        cudaGetLastError();  // not enough room for the estimate: let the pool grow on demand
                    }
                }
            }
            j->send_contig(tseq);  // in flight while the host walks the records
            j->timer.hbegin();
            records(*j);
            j->timer.hend("upload:host_parse");
            j->enqueue_arrays();   // copy stream
            j->timer.hbegin();
            j->send_seq();         // K0 gather on the (high-priority) copy stream
            j->timer.hend("upload:host_send_seq");
            NP2_CUDA(cudaStreamWaitEvent(ctx->stream, j->sc->ev_copied, 0));
            // spans and per-read arrays are on the device (in stream order): CIGAR words -> op records
            if (!j->ing.host_ops)
                cigar_ops(j->d_blob.p, j->d_seq_off.p, j->d_ncig.p, j->d_op_off.p, reinterpret_cast<uint4 *>(j->d_ops.p),
                          (uint32_t)j->ing.pos.size(), ctx->stream);
        } else {
            j->tseq.assign(tseq, tseq + tlen);
        }
        ctx->refs++;
        *out = j.release();
    });
}

int np2_job_create(np2_ctx *ctx, const uint8_t *tseq, uint32_t tlen, const uint8_t *bam, uint64_t bam_len,
                   np2_table *const *tables, uint32_t n_tables, const np2_opts *opts, np2_job **out) {
    return job_create_impl(
        ctx, tseq, tlen, bam_len,
        [&](np2_job &j) {
            j.bam = bam;
            j.bam_len = bam_len;
            parse_records(bam, bam_len, tlen, *opts, j.ing);
        },
        tables, n_tables, opts, out);
}

/* A contig straight from its BGZF members: they are inflated on the device and the records STAY there.  What the host
 * parse reads of a record — block_size, the 32 fixed bytes, the read name's length, the CIGAR words — is found and
 * gathered on the device (np2_inflate.cu: guessed record boundaries per 64 KiB chunk, walked per chunk, joined here) and
 * comes down as one compact buffer (~2 % of the record bytes); the SEQ + CIGAR spans are then gathered device to device.
 * The join accepts a chunk only when the chain that starts at the region's first byte ends exactly on the chunk's guessed
 * start; any miss, and any record the walk cannot take, sends the whole region through the host path instead (records
 * downloaded once, parse_records), so the result never depends on the guess. */
int np2_job_create_bgzf(np2_ctx *ctx, const uint8_t *tseq, uint32_t tlen, const uint8_t *comp, uint64_t comp_len,
                        const uint64_t *payload_off, const uint32_t *payload_len, const uint32_t *isize, uint32_t n_members,
                        uint64_t skip, uint64_t rec_len, np2_table *const *tables, uint32_t n_tables, const np2_opts *opts,
                        np2_job **out) {
    uint64_t hint = 0;
    for (uint32_t i = 0; isize && i < n_members; i++) hint += isize[i];
    return job_create_impl(
        ctx, tseq, tlen, hint,
        [&](np2_job &j) {
            cudaStream_t s = ctx->stream;
            uint64_t total = 0;
            bgzf_to_device(ctx, comp, comp_len, payload_off, payload_len, isize, n_members, j.d_records, total, nullptr);
            for (uint32_t i = 0; i < n_members; i++) j.h2d += payload_len[i];
            if (skip > total || rec_len > total - skip) throw np2::Error(NP2_ERR_ARG, "requested range outside the inflated members");
            const uint8_t *d_rec = j.d_records.p + skip;
            j.bam = rec_len ? d_rec : nullptr;
            j.bam_len = rec_len;
            j.records_on_device = true;
            if (!rec_len) {
                parse_heads(nullptr, nullptr, nullptr, 0, 0, tlen, *opts, j.ing);
                return;
            }
            auto host_path = [&]() {  // the speculation missed: the records come down once and are walked on the host
                std::vector<uint8_t> host(rec_len);
                NP2_CUDA(cudaMemcpyAsync(host.data(), d_rec, rec_len, cudaMemcpyDeviceToHost, s));
                NP2_CUDA(cudaStreamSynchronize(s));
                parse_records(host.data(), rec_len, tlen, *opts, j.ing);
                j.ing.n_fallback += 1u << 16;
            };
            const uint32_t nc = rec_chunk_count(rec_len);
            const uint64_t CH = rec_chunk_bytes();
            DBuf<uint64_t> d_start, d_end, d_hb, d_base;
            DBuf<uint32_t> d_cnt;
            d_start.alloc(nc, s);
            d_end.alloc(nc, s);
            d_hb.alloc(nc, s);
            d_cnt.alloc(nc, s);
            rec_chunk_starts(d_rec, rec_len, d_start.p, s);
            rec_chunk_count_walk(d_rec, rec_len, d_start.p, d_end.p, d_cnt.p, d_hb.p, s);
            std::vector<uint64_t> start(nc), end(nc), hb(nc), base(2 * (size_t)nc, ~0ull);
            std::vector<uint32_t> cnt(nc);
            d_start.download(start.data(), nc);
            d_end.download(end.data(), nc);
            d_hb.download(hb.data(), nc);
            d_cnt.download(cnt.data(), nc);
            NP2_CUDA(cudaStreamSynchronize(s));
            // join: follow the chain of chunks from byte 0
            uint64_t n_rec = 0, n_head = 0;
            bool ok = start[0] == 0;
            for (uint64_t c = 0; ok;) {
                base[c] = n_rec;
                base[nc + c] = n_head;
                n_rec += cnt[c];
                n_head += hb[c];
                const uint64_t e = end[c];
                if (e == ~0ull) ok = false;  // a record the walk could not take
                else if (e == rec_len) break;
                else if (e > rec_len || start[e / CH] != e) ok = false;
                else c = e / CH;
            }
            // NP2_BGZF_HOST_PARSE=1 (tests): take the fall-back although the join succeeded
            static const char *force_env = getenv("NP2_BGZF_HOST_PARSE");
            const bool force = force_env && atoi(force_env) != 0;
            if (!ok || force || n_rec >= (1ull << 31)) return host_path();
            DBuf<uint64_t> d_rec_off, d_head_off;
            DBuf<uint8_t> d_heads;
            d_base.alloc(2 * (size_t)nc, s);
            d_rec_off.alloc(std::max<uint64_t>(n_rec, 1), s);
            d_head_off.alloc(std::max<uint64_t>(n_rec, 1), s);
            d_heads.alloc(n_head + 64, s);
            NP2_CUDA(cudaMemcpyAsync(d_base.p, base.data(), 2 * (size_t)nc * 8, cudaMemcpyHostToDevice, s));
            rec_chunk_write_walk(d_rec, rec_len, d_start.p, d_base.p, d_base.p + nc, d_rec_off.p, d_head_off.p, d_heads.p, s);
            // heads | head offsets | record offsets -> one page-locked block of the context
            const size_t o_ho = (n_head + 64 + 15) & ~(size_t)15, o_ro = o_ho + n_rec * 8;
            ctx->p_infl_in.resize(o_ro + n_rec * 8 + 16);
            uint8_t *hp = ctx->p_infl_in.p;
            NP2_CUDA(cudaMemcpyAsync(hp, d_heads.p, n_head, cudaMemcpyDeviceToHost, s));
            NP2_CUDA(cudaMemcpyAsync(hp + o_ho, d_head_off.p, n_rec * 8, cudaMemcpyDeviceToHost, s));
            NP2_CUDA(cudaMemcpyAsync(hp + o_ro, d_rec_off.p, n_rec * 8, cudaMemcpyDeviceToHost, s));
            NP2_CUDA(cudaStreamSynchronize(s));
            j.d2h += n_head + n_rec * 16;
            parse_heads(hp, reinterpret_cast<const uint64_t *>(hp + o_ho), reinterpret_cast<const uint64_t *>(hp + o_ro), n_rec,
                        rec_len, tlen, *opts, j.ing);
        },
        tables, n_tables, opts, out);
}

/* ---- -S / --use_secondary (np2_secondary.cpp) */
struct np2_secmap {
    np2::SecMap *m;
};
int np2_secmap_create(np2_secmap **out) {
    return guard([&] {
        if (!out) throw np2::Error(NP2_ERR_ARG, "null argument");
        *out = new np2_secmap{np2::secmap_new()};
    });
}
void np2_secmap_destroy(np2_secmap *m) {
    if (!m) return;
    np2::secmap_delete(m->m);
    delete m;
}
int np2_secmap_scan_ids(np2_secmap *m, const uint8_t *bam, uint64_t bam_len) {
    return guard([&] {
        if (!m || (!bam && bam_len)) throw np2::Error(NP2_ERR_ARG, "null argument");
        np2::secmap_scan_ids(*m->m, bam, bam_len);
    });
}
int np2_secmap_scan_seqs(np2_secmap *m, const uint8_t *bam, uint64_t bam_len) {
    return guard([&] {
        if (!m || (!bam && bam_len)) throw np2::Error(NP2_ERR_ARG, "null argument");
        np2::secmap_scan_seqs(*m->m, bam, bam_len);
    });
}
int np2_secmap_fill(const np2_secmap *m, const uint8_t *bam, uint64_t bam_len, uint8_t *out, uint64_t cap, uint64_t *need) {
    return guard([&] {
        if (!m || (!bam && bam_len) || !need) throw np2::Error(NP2_ERR_ARG, "null argument");
        *need = np2::secmap_fill(*m->m, bam, bam_len, out, cap);
    });
}
uint64_t np2_secmap_size(const np2_secmap *m, uint64_t *n_seqs) { return m ? np2::secmap_counts(*m->m, n_seqs) : 0; }

void np2_set_host_threads(uint32_t n) { np2::set_host_threads(n); }
void np2_set_stage_timing(int on) { g_stage_events.store(on ? 1 : 0); }

/* BGZF members -> records, on the device (np2_inflate.cu).  The compressed span travels once (through a page-locked
 * ring when the caller's buffer is pageable, e.g. a memory-mapped file: host threads copy a chunk while the DMA moves the
 * previous one), a group of lanes inflates each member into its final place.  bgzf_to_device leaves the members' bytes,
 * back to back, in d_out (the stream is synchronised: the failure flag has been read). */
static void bgzf_to_device(np2_ctx *ctx, const uint8_t *comp, uint64_t comp_len, const uint64_t *payload_off,
                           const uint32_t *payload_len, const uint32_t *isize, uint32_t n_members, DBuf<uint8_t> &d_out,
                           uint64_t &total, float *kernel_ms) {
    if (!ctx || (n_members && (!comp || !payload_off || !payload_len || !isize))) throw np2::Error(NP2_ERR_ARG, "null argument");
    if (kernel_ms) *kernel_ms = 0;
    uint64_t lo = ~0ull, hi = 0;
    total = 0;
    for (uint32_t i = 0; i < n_members; i++) {
        if (payload_off[i] > comp_len || payload_len[i] > comp_len - payload_off[i])
            throw np2::Error(NP2_ERR_ARG, "BGZF member outside the compressed buffer");
        if (isize[i] > np2::infl::kMaxMember) throw np2::Error(NP2_ERR_FORMAT, "BAM/SAM parsing failed! (BGZF member larger than 64 KiB)");
        lo = std::min(lo, payload_off[i]);
        hi = std::max(hi, payload_off[i] + payload_len[i]);
        total += isize[i];
    }
    if (!n_members) return;
    NP2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const uint64_t span = hi - lo;
    constexpr uint64_t kFront = 16;  // the decoder reads whole aligned words: slack on both sides of the span
    DBuf<uint8_t> d_comp, d_args;
    d_comp.alloc(span + kFront + 32, s);
    d_out.alloc(total + 64, s);
    // member table: off | out_off | clen | isize | bad[2], one pinned block
    const size_t o_out = (size_t)n_members * 8, o_clen = o_out + (size_t)n_members * 8, o_isz = o_clen + (size_t)n_members * 4;
    const size_t o_bad = o_isz + (size_t)n_members * 4, args_bytes = o_bad + 16;
    ctx->p_infl_args.resize(args_bytes + 16);
    uint8_t *pa = ctx->p_infl_args.p;
    {
        uint64_t *a_off = reinterpret_cast<uint64_t *>(pa), *a_out = reinterpret_cast<uint64_t *>(pa + o_out);
        uint32_t *a_clen = reinterpret_cast<uint32_t *>(pa + o_clen), *a_isz = reinterpret_cast<uint32_t *>(pa + o_isz);
        uint32_t *a_bad = reinterpret_cast<uint32_t *>(pa + o_bad);
        uint64_t w = 0;
        for (uint32_t i = 0; i < n_members; i++) {
            a_off[i] = kFront + payload_off[i] - lo;
            a_out[i] = w;
            a_clen[i] = payload_len[i];
            a_isz[i] = isize[i];
            w += isize[i];
        }
        a_bad[0] = 0;
        a_bad[1] = 0xFFFFFFFFu;
    }
    d_args.alloc(args_bytes, s);
    NP2_CUDA(cudaMemcpyAsync(d_args.p, pa, args_bytes, cudaMemcpyHostToDevice, s));
    NP2_CUDA(cudaMemsetAsync(d_comp.p, 0, kFront, s));
    NP2_CUDA(cudaMemsetAsync(d_comp.p + kFront + span, 0, 32, s));
    bool pinned = false;
    {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, comp) != cudaSuccess) cudaGetLastError();
        else pinned = at.type == cudaMemoryTypeHost;
    }
    if (pinned) {
        NP2_CUDA(cudaMemcpyAsync(d_comp.p + kFront, comp + lo, span, cudaMemcpyHostToDevice, s));
    } else {
        const uint64_t kChunk = 8ull << 20;
        const uint64_t n_chunks = (span + kChunk - 1) / kChunk;
        ctx->p_infl_in.resize(std::min<uint64_t>(span, 2 * kChunk));  // two halves: copy into one while the other is on the link
        cudaEvent_t ev[2] = {nullptr, nullptr};
        NP2_CUDA(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
        NP2_CUDA(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
        const unsigned T = std::max(1u, np2::host_threads());
        try {
            for (uint64_t c = 0; c < n_chunks; c++) {
                const uint64_t b = c * kChunk, e = std::min(span, b + kChunk);
                uint8_t *half = ctx->p_infl_in.p + (c & 1) * kChunk;
                if (c >= 2) NP2_CUDA(cudaEventSynchronize(ev[c & 1]));
                np2::parallel_for(T, [&](unsigned ti) {
                    const uint64_t tb = b + (e - b) * ti / T, te = b + (e - b) * (ti + 1) / T;
                    np2::copy_streaming(half + (tb - b), comp + lo + tb, te - tb);
                    np2::store_fence();
                });
                NP2_CUDA(cudaMemcpyAsync(d_comp.p + kFront + b, half, e - b, cudaMemcpyHostToDevice, s));
                NP2_CUDA(cudaEventRecord(ev[c & 1], s));
            }
        } catch (...) {
            cudaStreamSynchronize(s);
            cudaEventDestroy(ev[0]);
            cudaEventDestroy(ev[1]);
            throw;
        }
        NP2_CUDA(cudaStreamSynchronize(s));  // the ring is reused by the next call
        cudaEventDestroy(ev[0]);
        cudaEventDestroy(ev[1]);
    }
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    if (kernel_ms) {
        NP2_CUDA(cudaEventCreate(&t0));
        NP2_CUDA(cudaEventCreate(&t1));
        NP2_CUDA(cudaEventRecord(t0, s));
    }
    uint32_t *d_bad = reinterpret_cast<uint32_t *>(d_args.p + o_bad);
    bgzf_inflate(d_comp.p, reinterpret_cast<const uint64_t *>(d_args.p), reinterpret_cast<const uint32_t *>(d_args.p + o_clen),
                 reinterpret_cast<const uint64_t *>(d_args.p + o_out), reinterpret_cast<const uint32_t *>(d_args.p + o_isz),
                 n_members, d_out.p, d_bad, s);
    if (kernel_ms) NP2_CUDA(cudaEventRecord(t1, s));
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) throw np2::Error(NP2_ERR_CUDA, std::string("k_bgzf_inflate: ") + cudaGetErrorString(le));
    uint32_t *h_bad = reinterpret_cast<uint32_t *>(pa + args_bytes);  // behind the uploaded part of the block
    NP2_CUDA(cudaMemcpyAsync(h_bad, d_bad, 8, cudaMemcpyDeviceToHost, s));
    NP2_CUDA(cudaStreamSynchronize(s));
    if (kernel_ms) {
        cudaEventElapsedTime(kernel_ms, t0, t1);
        cudaEventDestroy(t0);
        cudaEventDestroy(t1);
    }
    if (h_bad[0])
        throw np2::Error(NP2_ERR_FORMAT, "BAM/SAM parsing failed! (BGZF member " + std::to_string(h_bad[1]) + " of " +
                                             std::to_string(n_members) + " does not inflate to its ISIZE)");
}

int np2_bgzf_inflate(np2_ctx *ctx, const uint8_t *comp, uint64_t comp_len, const uint64_t *payload_off,
                     const uint32_t *payload_len, const uint32_t *isize, uint32_t n_members, uint64_t skip, uint64_t out_len,
                     uint8_t *out, float *kernel_ms) {
    return guard([&] {
        if (out_len && !out) throw np2::Error(NP2_ERR_ARG, "null argument");
        DBuf<uint8_t> d_out;
        uint64_t total = 0;
        bgzf_to_device(ctx, comp, comp_len, payload_off, payload_len, isize, n_members, d_out, total, kernel_ms);
        if (skip > total || out_len > total - skip) throw np2::Error(NP2_ERR_ARG, "requested range outside the inflated members");
        if (!out_len) return;
        NP2_CUDA(cudaMemcpyAsync(out, d_out.p + skip, out_len, cudaMemcpyDeviceToHost, ctx->stream));
        NP2_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}

int np2_host_alloc(uint64_t bytes, void **out) {
    return guard([&] {
        if (!out) throw np2::Error(NP2_ERR_ARG, "null argument");
        NP2_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable | cudaHostAllocMapped));
    });
}
void np2_host_free(void *p) {
    if (p) cudaFreeHost(p);
}
int np2_job_ingest_path(const np2_job *job) { return job->seq_path; }

int np2_job_upload(np2_job *job) {
    return guard([&] {
        NP2_CUDA(cudaSetDevice(job->ctx->device));
        if (job->L >= job->opt.min_ctg_len) job->upload();
    });
}

int np2_job_run(np2_job *job, int32_t dump_iter) {
    return guard([&] {
        NP2_CUDA(cudaSetDevice(job->ctx->device));
        job->run(dump_iter);
    });
}

void np2_job_destroy(np2_job *job) {
    if (!job) return;
    cudaSetDevice(job->ctx->device);
    cudaStreamSynchronize(job->ctx->stream);
    np2_ctx *c = job->ctx;
    delete job;
    ctx_release(c);
}

int np2_polish_contig(np2_ctx *ctx, const uint8_t *tseq, uint32_t tlen, const uint8_t *bam, uint64_t bam_len,
                      np2_table *const *tables, uint32_t n_tables, const np2_opts *opts, np2_job **out) {
    np2_job *j = nullptr;
    int rc = np2_job_create(ctx, tseq, tlen, bam, bam_len, tables, n_tables, opts, &j);
    if (rc == NP2_OK) rc = np2_job_upload(j);
    if (rc == NP2_OK) rc = np2_job_run(j, -1);
    if (rc != NP2_OK) {
        np2_job_destroy(j);
        return rc;
    }
    *out = j;
    return NP2_OK;
}

uint64_t np2_job_get_consensus(np2_job *j, const uint32_t **pos, const uint8_t **base) {
    if (pos) {  // ConsensusBase.pos is materialised on request; the FASTA record only needs first/last
        if (!j->res_pos_valid) {
            cudaSetDevice(j->ctx->device);
            j->p_cpos.resize(std::max(j->res_N, 1u));
            cudaMemcpyAsync(j->p_cpos.p, j->jd_cpos.p, (size_t)j->res_N * 4, cudaMemcpyDeviceToHost, j->ctx->stream);
            cudaStreamSynchronize(j->ctx->stream);
            if (j->res_sparse) {  // rebuild the full region list from the device copies
                const uint32_t nr = j->res_nreg;
                std::vector<uint32_t> st(nr), ra(nr), rb(nr), sl(nr);
                cudaMemcpyAsync(st.data(), j->jd_reg_start.p, (size_t)nr * 4, cudaMemcpyDeviceToHost, j->ctx->stream);
                cudaMemcpyAsync(ra.data(), j->jd_reg_a.p, (size_t)nr * 4, cudaMemcpyDeviceToHost, j->ctx->stream);
                cudaMemcpyAsync(rb.data(), j->jd_reg_b.p, (size_t)nr * 4, cudaMemcpyDeviceToHost, j->ctx->stream);
                cudaMemcpyAsync(sl.data(), j->jd_reg_len.p, (size_t)nr * 4, cudaMemcpyDeviceToHost, j->ctx->stream);
                cudaStreamSynchronize(j->ctx->stream);
                Patched &pc = j->res_patch;
                pc.reset(nr);
                for (uint32_t r = 0; r < nr; r++) {
                    const uint32_t q = nr - 1 - r;
                    pc.start[q] = st[r];
                    pc.a[q] = ra[r];
                    pc.b[q] = rb[r];
                    pc.seed[q].len = sl[r];
                }
                j->res_sparse = false;
            }
            j->res_patch.cpos = j->p_cpos.p;
            j->res_patch.N = j->res_N;
            positions(j->res_patch, j->res_pos);
            j->res_pos_valid = true;
        }
        *pos = j->res_pos.data();
    }
    if (base) *base = j->res_base.p;
    return j->res_base.n;
}
uint64_t np2_job_get_span(np2_job *j, uint32_t *first_pos, uint32_t *last_pos) {
    *first_pos = j->res_first;
    *last_pos = j->res_last;
    return j->res_base.n;
}
uint64_t np2_job_get_reads(np2_job *j, const int32_t **rec_idx, const uint32_t **t_s, const uint32_t **t_e,
                           const uint64_t **nib_off, const uint8_t **nib, const uint8_t **blank) {
    *rec_idx = j->d_rec_idx.data();
    *t_s = j->dm_ts.data();
    *t_e = j->dm_te.data();
    *nib_off = j->dm_nib_off.data();
    *nib = j->dm_nib.data();
    *blank = j->dm_blank.data();
    return j->d_rec_idx.size();
}
uint64_t np2_job_get_msa(np2_job *j, const uint64_t **off, const uint16_t **bases, const uint16_t **delta,
                         const uint32_t **count, const uint32_t **besti) {
    *off = j->dm_msa_off.data();
    *bases = j->dm_msa_bases.data();
    *delta = j->dm_msa_delta.data();
    *count = j->dm_msa_count.data();
    *besti = j->dm_msa_besti.data();
    return j->dm_msa_bases.size();
}
uint64_t np2_job_get_dp_consensus(np2_job *j, const uint32_t **pos, const uint8_t **base, const uint8_t **flags) {
    *pos = j->dm_dp_pos.data();
    *base = j->dm_dp_base.data();
    *flags = j->dm_dp_flags.data();
    return j->dm_dp_pos.size();
}
uint64_t np2_job_get_regions(np2_job *j, const uint32_t **start, const uint32_t **end, const uint8_t **lable) {
    *start = j->dm_reg_start.data();
    *end = j->dm_reg_end.data();
    *lable = j->dm_reg_lable.data();
    return j->dm_reg_start.size();
}
uint64_t np2_job_get_candidates(np2_job *j, const uint64_t **roff, const uint32_t **order, const uint16_t **kscore,
                                const uint64_t **kmer, const uint64_t **seq_off, const uint8_t **seq) {
    *roff = j->dm_can_roff.data();
    *order = j->dm_can_order.data();
    *kscore = j->dm_can_kscore.data();
    *kmer = j->dm_can_kmer.data();
    *seq_off = j->dm_can_seq_off.data();
    *seq = j->dm_can_seq.data();
    return j->dm_can_order.size();
}
uint64_t np2_job_get_dropped(np2_job *j, const uint32_t **ids) {
    *ids = j->dm_dropped.data();
    return j->dm_dropped.size();
}
uint64_t np2_job_get_pair_weights(np2_job *j, const uint64_t **keys, const int64_t **vals) {
    *keys = j->dm_pair_key.data();
    *vals = j->dm_pair_val.data();
    return j->dm_pair_key.size();
}

uint32_t np2_job_get_timings(np2_job *j, const char **names, const float **ms, const uint32_t **launches) {
    j->timing_names.clear();
    for (auto &n : j->timer.names) {
        j->timing_names += n;
        j->timing_names.push_back('\0');
    }
    *names = j->timing_names.data();
    *ms = j->timer.ms.data();
    *launches = j->timer.launches.data();
    return (uint32_t)j->timer.names.size();
}
void np2_job_get_traffic(np2_job *j, uint64_t *h2d_bytes, uint64_t *d2h_bytes, uint64_t *n_kernel_launches,
                         uint64_t *n_alignment_columns, uint64_t *n_probes) {
    if (h2d_bytes) *h2d_bytes = j->h2d;
    if (d2h_bytes) *d2h_bytes = j->d2h;
    if (n_kernel_launches) *n_kernel_launches = j->n_launch;
    if (n_alignment_columns) *n_alignment_columns = j->ing.total_cols;
    if (n_probes) *n_probes = j->n_probes;
}

void np2_job_get_stats(np2_job *j, uint64_t out[12]) {
    memcpy(out, j->stats, sizeof j->stats);
    out[8] = j->n_spec_ok;
    out[9] = j->n_respec;
    out[10] = j->n_sync;
    out[11] = 0;
}

int np2_debug_phase(const uint64_t *keys, const int64_t *vals, uint64_t n_edges, uint32_t model, uint32_t use_all_reads,
                    uint32_t *dropped, uint64_t cap, uint64_t *n_dropped, uint32_t *path) {
    return guard([&] {
        if ((!keys || !vals) && n_edges) throw np2::Error(NP2_ERR_ARG, "null argument");
        std::vector<uint32_t> d = phase_reads(keys, reinterpret_cast<const long long *>(vals), n_edges, model == 0,
                                              use_all_reads != 0);
        if (path) *path = (uint32_t)np2::phase_last_path();
        if (n_dropped) *n_dropped = d.size();
        if (dropped) memcpy(dropped, d.data(), std::min<uint64_t>(cap, d.size()) * 4);
    });
}

int np2_debug_parse(const uint8_t *bam, uint64_t bam_len, uint32_t tlen, const np2_opts *opts, uint32_t threads,
                    uint64_t out[6]) {
    return guard([&] {
        Ingest ing;
        // threads bit 16: the job path's parse (CIGAR summed, no op records); bit 17: op records built but left out of
        // the digest (so that the two can be compared)
        // bit 18: the parse np2_job_create_bgzf runs — only the heads of the records (gathered here by a plain walk) and
        // their offsets are looked at; digest of the per-record arrays only
        const bool fast = threads & 0x10000u, heads_only = threads & 0x40000u, scalars_only = threads & 0x70000u;
        if (heads_only) {
            std::vector<uint8_t> heads;
            std::vector<uint64_t> head_off, rec_off;
            for (uint64_t p = 0; p < bam_len;) {
                if (p + 36 > bam_len) throw np2::Error(NP2_ERR_FORMAT, "BAM/SAM parsing failed!");
                int32_t bs;
                memcpy(&bs, bam + p, 4);
                uint16_t n_cig;
                memcpy(&n_cig, bam + p + 16, 2);
                const uint64_t head = 36ull + bam[p + 12] + 4ull * n_cig;
                if (bs < 32 || p + 4 + (uint64_t)bs > bam_len || head > 4ull + (uint64_t)bs)
                    throw np2::Error(NP2_ERR_FORMAT, "BAM/SAM parsing failed!");
                rec_off.push_back(p);
                head_off.push_back(heads.size());
                heads.insert(heads.end(), bam + p, bam + p + head);
                p += 4 + (uint64_t)bs;
            }
            heads.resize(heads.size() + 64);
            parse_heads(heads.data(), head_off.data(), rec_off.data(), rec_off.size(), bam_len, tlen, *opts, ing, threads & 0xFFFFu);
        } else {
            parse_records(bam, bam_len, tlen, *opts, ing, threads & 0xFFFFu, !fast);
        }
        uint64_t h = 0xcbf29ce484222325ull;
        auto mix = [&](const void *p, size_t n) {
            const uint8_t *b = static_cast<const uint8_t *>(p);
            for (size_t i = 0; i < n; i++) h = (h ^ b[i]) * 0x100000001b3ull;
        };
        auto mixv = [&](const auto &v) { mix(v.data(), v.size() * sizeof(v[0])); };
        mixv(ing.all_tid);
        mixv(ing.all_pos);
        mixv(ing.rec_idx);
        mixv(ing.pos);
        mixv(ing.ncols);
        mixv(ing.rlen);
        mixv(ing.rspan);
        mixv(ing.is_clip);
        mixv(ing.seq_off);
        mixv(ing.seq_bytes);
        mixv(ing.op_off);
        mixv(ing.nib_off);
        mixv(ing.ck_off);
        mixv(ing.n_cig);
        for (int a = 0; a < 4 && !scalars_only; a++)
            for (const Ingest::OpChunk &c : ing.op_chunks)
                for (size_t i = 0; i < c.n; i++) {
                    const uint32_t v = a == 0 ? c.ops[i].col : a == 1 ? c.ops[i].q : a == 2 ? c.ops[i].t : c.ops[i].cig;
                    mix(&v, 4);
                }
        out[0] = ing.all_tid.size();
        out[1] = ing.pos.size();
        out[2] = ing.n_ops;
        out[3] = ing.total_cols;
        out[4] = ing.n_fallback;
        out[5] = h;
    });
}

uint64_t np2_format_fasta(const char *tid, const uint32_t *pos, const uint8_t *base, uint64_t n, int uppercase,
                          int out_pos, uint8_t *out, uint64_t cap) {
    // display_consensusbase_vec main.rs:607-645
    auto up = [&](uint8_t c) -> uint8_t { return (uppercase && c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c; };
    uint64_t w = 0;
    auto put = [&](const char *s, size_t len) {
        if (w + len <= cap) memcpy(out + w, s, len);
        w += len;
    };
    char num[32];
    const size_t tl = strlen(tid);
    if (out_pos) {
        for (uint64_t i = 0; i < n; i++) {
            put(tid, tl);
            char b[3] = {'\t', (char)up(base[i]), '\t'};
            put(b, 3);
            int l = snprintf(num, sizeof num, "%u\n", pos[i]);
            put(num, (size_t)l);
        }
    } else if (n) {
        put(">", 1);
        put(tid, tl);
        int l = snprintf(num, sizeof num, " start:%u", pos[0]);
        put(num, (size_t)l);
        l = snprintf(num, sizeof num, " end:%u\n", pos[n - 1]);
        put(num, (size_t)l);
        if (w + n <= cap) {
            if (uppercase)
                for (uint64_t i = 0; i < n; i++) out[w + i] = up(base[i]);
            else
                memcpy(out + w, base, n);
        }
        w += n;
        put("\n", 1);
    }
    return w;
}

}  // extern "C"
