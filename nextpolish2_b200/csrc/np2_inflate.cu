// np2_inflate.cu — BGZF members inflated on the device (replaces htslib's bgzf_read_block +
// zlib under bam::IndexedReader::fetch / records(), reference src/main.rs:1745-1757).
//
// A coordinate-sorted HiFi BAM holds ~700 independent <= 64 KiB DEFLATE members per Mbp of contig at 30x.  One WARP
// takes one member: lane 0 runs the bit-serial part (np2_inflate.cuh infl_step: block headers, Huffman tables in shared
// memory, symbol decode, literal bytes), and every time bytes have to be COPIED — an LZ77 match or a stored block — all 32
// lanes do it, coalesced.  The serial part of a warp is latency bound (table look-up -> shift -> look-up), so the SM hides
// it behind the other resident warps (one member each): throughput comes from the ~6000 members in flight, not from
// any single one.  Output goes straight to its final place in the contig's record buffer (members are laid out back to
// back by the caller's prefix sum of ISIZE), compressed input is read through the read-only path.
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
    static bool attr_set = false;
    if (!attr_set) {
        NP2_CUDA(cudaFuncSetAttribute(k_bgzf_inflate<G, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
        attr_set = true;
    }
    NP2_K((k_bgzf_inflate<G, MINB>))<<<(n_members + kPerCta - 1) / kPerCta, kInflThreads, kSmem, s>>>(
        d_comp, d_off, d_clen, d_out_off, d_isize, n_members, d_out, d_bad, inline_max);
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

}  // namespace np2
