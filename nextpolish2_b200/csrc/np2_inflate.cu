// np2_inflate.cu — BGZF members inflated on the device (replaces htslib's bgzf_read_block +
// zlib under bam::IndexedReader::fetch / records(), reference src/main.rs:1745-1757).
//
// A coordinate-sorted HiFi BAM holds ~900 independent <= 64 KiB DEFLATE members per Mbp of contig at 30x.  A GROUP OF
// LANES (16 by default, two members per warp) takes one member: the group's leader runs the bit-serial part
// (np2_inflate.cuh infl_step: block headers, Huffman tables in shared memory, symbol decode, literal bytes, matches of up
// to 8 bytes), and every time more bytes have to be COPIED — a longer LZ77 match or a stored block — the whole group does
// it, coalesced.  One member's decode is a serial chain (table look-up -> shift -> look-up); the SM hides it behind the
// other resident groups: throughput comes from the ~9000 members in flight, not from any single one, and once the grid
// fills the SMs the kernel is bound by instruction issue (profiles/r02aw_inflate_ncu_summary.txt).  Output goes straight
// to its final place in the contig's record buffer (members are laid out back to back by the caller's prefix sum of
// ISIZE), compressed input is read through the read-only path.  The second half of the file finds the record boundaries
// in the inflated bytes, so that the records can stay on the device (np2_job_create_bgzf).
#include <algorithm>
#include <cstdlib>

#include "np2_common.cuh"
#include "np2_inflate.cuh"
#include "np2_kernels.cuh"

namespace np2 {

namespace {
constexpr uint32_t kInflThreads = 256;

// G lanes per member (8, 16 or 32): the leader lane of a group decodes, the group copies.  With G < 32 a warp carries
// 32 / G independent decode chains at the register cost of one, which is what hides the chains' latency.
template <uint32_t G, int MINB>
__global__ void __launch_bounds__(kInflThreads, MINB) k_bgzf_inflate(const uint8_t *__restrict__ comp,
                                                               const uint64_t *__restrict__ m_off,
                                                               const uint32_t *__restrict__ m_clen,
                                                               const uint64_t *__restrict__ m_out,
                                                               const uint32_t *__restrict__ m_isize, uint32_t n,
                                                               uint8_t *__restrict__ out, uint32_t *__restrict__ bad,
                                                               uint32_t inline_max) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    infl::Tabs *tabs = reinterpret_cast<infl::Tabs *>(smem_raw);
    constexpr uint32_t kPerCta = kInflThreads / G;
    const uint32_t grp = threadIdx.x / G, gl = threadIdx.x % G;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t mask = G == 32 ? 0xFFFFFFFFu : ((1u << G) - 1u) << (lane - gl);
    const int leader = (int)(lane - gl);
    const uint32_t m = blockIdx.x * kPerCta + grp;
    if (m >= n) return;  // whole groups leave together
    const uint8_t *payload = comp + m_off[m];
    uint8_t *o = out + m_out[m];
    const uint32_t isize = m_isize[m];
    infl::State st;
    if (gl == 0) infl::init(st, payload, m_clen[m], o, isize, inline_max);
    uint32_t pos = 0;
    bool ok = false;
    for (;;) {
        uint32_t ev = infl::EV_ERROR, a = 0, b = 0;
        if (gl == 0) {
            ev = infl::infl_step(st, tabs[grp], a, b);
            pos = st.pos;
        }
        ev = __shfl_sync(mask, ev, leader);
        pos = __shfl_sync(mask, pos, leader);
        if (ev == infl::EV_DONE) {
            ok = pos == isize;
            break;
        }
        if (ev == infl::EV_ERROR) break;
        a = __shfl_sync(mask, a, leader);
        b = __shfl_sync(mask, b, leader);
        __syncwarp(mask);  // the leader's literal stores are visible to the lanes that copy from them
        uint8_t *d = o + pos;
        if (ev == infl::EV_STORED) {
            const uint8_t *src = payload + b;
            for (uint32_t i = gl; i < a; i += G) d[i] = __ldg(src + i);
        } else if (b >= a) {  // source and destination do not overlap
            const uint8_t *src = d - b;
            for (uint32_t i = gl; i < a; i += G) d[i] = src[i];
        } else if (b == 1) {  // run of one byte
            const uint8_t v = d[-1];
            for (uint32_t i = gl; i < a; i += G) d[i] = v;
        } else {  // the last b bytes repeat: byte i of the match is byte i mod b of that period
            const uint8_t *src = d - b;
            for (uint32_t i = gl; i < a; i += G) d[i] = src[i % b];
        }
        __syncwarp(mask);
        if (gl == 0) st.pos = pos + a;
    }
    if (!ok && gl == 0) {
        atomicAdd(bad, 1u);
        atomicMin(bad + 1, m);
    }
}

template <uint32_t G, int MINB>
void launch_inflate(const uint8_t *d_comp, const uint64_t *d_off, const uint32_t *d_clen, const uint64_t *d_out_off,
                    const uint32_t *d_isize, uint32_t n_members, uint8_t *d_out, uint32_t *d_bad, uint32_t inline_max,
                    cudaStream_t s) {
    constexpr uint32_t kPerCta = kInflThreads / G;
    constexpr int kSmem = (int)(kPerCta * sizeof(infl::Tabs));
    // above the 48 KB every kernel may use without asking (8 lanes per member: 32 tables per CTA); the attribute belongs
    // to the current device, so it is set at every launch rather than once per process
    if (kSmem > 48 * 1024) NP2_CUDA(cudaFuncSetAttribute(k_bgzf_inflate<G, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    NP2_K((k_bgzf_inflate<G, MINB>))<<<(n_members + kPerCta - 1) / kPerCta, kInflThreads, kSmem, s>>>(
        d_comp, d_off, d_clen, d_out_off, d_isize, n_members, d_out, d_bad, inline_max);
}
/* ------------------------------------------------------------------ record boundaries of an inflated contig */
// The records of a contig are a linked list through hundreds of MB (every record starts with its own length).  The
// host parser follows it from byte ranges it guesses a record boundary in (np2_host.cpp parse_records); the same
// scheme runs here so that the inflated records never have to leave the device: the region is cut into 64 KiB chunks,
// a warp finds the first position of its chunk where four plausible record headers follow each other, a thread walks
// the chunk's records from there, and the HOST joins the chunks (a chunk counts only if the chain that starts at byte 0
// lands exactly on its guessed start: np2_api.cu).  A second walk then writes every record's offset and gathers its
// head — block_size, the 32 fixed bytes, the read name and the CIGAR words, all the host parse ever reads — into one
// compact buffer that crosses the link instead of the records.
constexpr uint32_t kRecChunk = 65536;
constexpr uint64_t kNoStart = ~0ull;

__device__ __forceinline__ uint32_t ld32u(const uint8_t *p) {
    return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24;
}
// the host's plausibility rule (np2_host.cpp) + the record must belong to this reference
__device__ __forceinline__ bool rec_plausible(const uint8_t *R, uint64_t n, uint64_t p, int32_t tid, uint64_t &next) {
    if (p + 36 > n) return false;
    if ((int32_t)ld32u(R + p + 4) != tid) return false;
    const int32_t bs = (int32_t)ld32u(R + p);
    if (bs < 32 || p + 4 + (uint64_t)bs > n) return false;
    const int32_t pos = (int32_t)ld32u(R + p + 8), l_seq = (int32_t)ld32u(R + p + 20);
    const uint32_t l_name = R[p + 12], n_cig = (uint32_t)R[p + 16] | (uint32_t)R[p + 17] << 8;
    if (pos < -1 || l_seq < 0 || l_name == 0) return false;
    if (32ull + l_name + 4ull * n_cig + ((uint64_t)l_seq + 1) / 2 > (uint64_t)bs) return false;
    if (R[p + 4 + 32 + l_name - 1] != 0) return false;
    next = p + 4 + (uint64_t)bs;
    return true;
}
__global__ void __launch_bounds__(128) k_rec_chunk_start(const uint8_t *__restrict__ R, uint64_t n,
                                                         uint64_t *__restrict__ start, uint32_t n_chunks) {
    const uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (c >= n_chunks) return;
    const int32_t tid = n >= 8 ? (int32_t)ld32u(R + 4) : -1;  // refID of the first record: the region is one reference's
    if (c == 0) {  // the region begins with a record
        if (lane == 0) start[0] = 0;
        return;
    }
    const uint64_t lo = (uint64_t)c * kRecChunk, hi = min(n, lo + kRecChunk);
    for (uint64_t base = lo; base < hi; base += 32) {
        const uint64_t p = base + lane;
        bool ok = p < hi;
        uint64_t x = p;
        for (int k = 0; ok && k < 4; k++) {
            if (x == n) break;  // the chain reaches the end of the region
            uint64_t nx;
            ok = rec_plausible(R, n, x, tid, nx);
            x = nx;
        }
        const uint32_t hit = __ballot_sync(0xFFFFFFFFu, ok);
        if (hit) {
            if (lane == 0) start[c] = base + (uint32_t)(__ffs((int)hit) - 1);
            return;
        }
    }
    if (lane == 0) start[c] = kNoStart;
}
// WRITE = false: per chunk the number of records that start in it, the bytes of their heads and where the walk ends.
// WRITE = true : the same walk for the chunks the host accepted (rec_base / head_base = exclusive prefix sums over them,
//                ~0 for a chunk that is not on the chain): record offsets, head offsets, the head bytes themselves.
template <bool WRITE>
__global__ void __launch_bounds__(128) k_rec_chunk_walk(const uint8_t *__restrict__ R, uint64_t n,
                                                        const uint64_t *__restrict__ start, uint32_t n_chunks,
                                                        uint64_t *__restrict__ end, uint32_t *__restrict__ cnt,
                                                        uint64_t *__restrict__ hbytes, const uint64_t *__restrict__ rec_base,
                                                        const uint64_t *__restrict__ head_base, uint64_t *__restrict__ rec_off,
                                                        uint64_t *__restrict__ head_off, uint8_t *__restrict__ heads) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    uint64_t p = start[c];
    if (WRITE && (p == kNoStart || rec_base[c] == ~0ull)) return;
    if (!WRITE && p == kNoStart) {
        end[c] = kNoStart;
        cnt[c] = 0;
        hbytes[c] = 0;
        return;
    }
    const uint64_t hi = min(n, ((uint64_t)c + 1) * kRecChunk);
    uint32_t k = 0;
    uint64_t hb = 0;
    while (p < hi) {
        if (p + 36 > n) {  // a truncated record: the host reports it
            p = kNoStart;
            break;
        }
        const int32_t bs = (int32_t)ld32u(R + p);
        const uint32_t l_name = R[p + 12], n_cig = (uint32_t)R[p + 16] | (uint32_t)R[p + 17] << 8;
        const uint64_t head = 36ull + l_name + 4ull * n_cig;
        if (bs < 32 || p + 4 + (uint64_t)bs > n || head > 4ull + (uint64_t)bs) {
            p = kNoStart;
            break;
        }
        if (WRITE) {
            const uint64_t i = rec_base[c] + k, ho = head_base[c] + hb;
            rec_off[i] = p;
            head_off[i] = ho;
            for (uint64_t b = 0; b < head; b++) heads[ho + b] = R[p + b];
        }
        k++;
        hb += head;
        p += 4 + (uint64_t)bs;
    }
    if (!WRITE) {
        end[c] = p;
        cnt[c] = k;
        hbytes[c] = hb;
    }
}

int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}
}  // namespace

// d_bad: two words, {0, 0xFFFFFFFF} before the launch -> {members that failed, index of the first one}
// Tuning knobs (read at every call; profiles/bgzf_inflate_bench.py sweeps them): NP2_INFLATE_LANES = lanes per member
// (8 / 16 / 32), NP2_INFLATE_INLINE = longest match the decoding lane copies itself (0..8), NP2_INFLATE_MINB = resident
// CTAs per SM the kernel is compiled for (3: 80 registers, 4: 64 registers with a few spills).
void bgzf_inflate(const uint8_t *d_comp, const uint64_t *d_off, const uint32_t *d_clen, const uint64_t *d_out_off,
                  const uint32_t *d_isize, uint32_t n_members, uint8_t *d_out, uint32_t *d_bad, cudaStream_t s) {
    if (!n_members) return;
    const int lanes = env_int("NP2_INFLATE_LANES", 16), minb = env_int("NP2_INFLATE_MINB", 4);
    const uint32_t inl = (uint32_t)std::max(0, env_int("NP2_INFLATE_INLINE", (int)infl::kInlineMatch));
#define NP2_INFL(G, MB) launch_inflate<G, MB>(d_comp, d_off, d_clen, d_out_off, d_isize, n_members, d_out, d_bad, inl, s)
    if (lanes == 32) minb == 3 ? NP2_INFL(32, 3) : NP2_INFL(32, 4);
    else if (lanes == 8) minb == 3 ? NP2_INFL(8, 3) : NP2_INFL(8, 4);
    else minb == 3 ? NP2_INFL(16, 3) : NP2_INFL(16, 4);
#undef NP2_INFL
}

uint32_t rec_chunk_count(uint64_t n) { return (uint32_t)((n + kRecChunk - 1) / kRecChunk); }
uint32_t rec_chunk_bytes() { return kRecChunk; }
void rec_chunk_starts(const uint8_t *d_rec, uint64_t n, uint64_t *d_start, cudaStream_t s) {
    const uint32_t nc = rec_chunk_count(n);
    if (nc) NP2_K(k_rec_chunk_start)<<<(nc * 32 + 127) / 128, 128, 0, s>>>(d_rec, n, d_start, nc);
}
void rec_chunk_count_walk(const uint8_t *d_rec, uint64_t n, const uint64_t *d_start, uint64_t *d_end, uint32_t *d_cnt,
                          uint64_t *d_hbytes, cudaStream_t s) {
    const uint32_t nc = rec_chunk_count(n);
    if (nc)
        NP2_K(k_rec_chunk_walk<false>)<<<(nc + 127) / 128, 128, 0, s>>>(d_rec, n, d_start, nc, d_end, d_cnt, d_hbytes, nullptr,
                                                                        nullptr, nullptr, nullptr, nullptr);
}
void rec_chunk_write_walk(const uint8_t *d_rec, uint64_t n, const uint64_t *d_start, const uint64_t *d_rec_base,
                          const uint64_t *d_head_base, uint64_t *d_rec_off, uint64_t *d_head_off, uint8_t *d_heads,
                          cudaStream_t s) {
    const uint32_t nc = rec_chunk_count(n);
    if (nc)
        NP2_K(k_rec_chunk_walk<true>)<<<(nc + 127) / 128, 128, 0, s>>>(d_rec, n, d_start, nc, nullptr, nullptr, nullptr, d_rec_base,
                                                                       d_head_base, d_rec_off, d_head_off, d_heads);
}

}  // namespace np2
