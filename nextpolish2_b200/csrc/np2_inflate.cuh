// np2_inflate.cuh — raw DEFLATE (RFC 1951) decoder for BGZF members, written so that the SAME source runs as the lane-0
// half of the device kernel (np2_inflate.cu: one warp per BGZF member) and as a plain host function (the non-GPU tests
// compile this header with g++ and compare every member with zlib).
//
// What it replaces: the BGZF layer under rust-htslib's bam::IndexedReader (reference src/main.rs:1745-1757: fetch +
// records()), i.e. htslib's bgzf_read_block -> zlib inflate of each <= 64 KiB member.  A BGZF member is
// an independent DEFLATE stream without preset dictionary, so members decode in parallel and a match never reaches
// before the member's own output.
//
// Split of the work (event protocol): `infl_step` reads block headers, builds the Huffman tables and decodes symbols,
// writing literals and SHORT matches itself; whenever more bytes have to be COPIED (a longer LZ77 match, or the payload of
// a stored block) it returns an event and the caller copies — the member's group of lanes on the device, a loop on the
// host — then calls again.
//
// Tables (per member, shared memory on the device): 9-bit / 7-bit first-level lookup for the literal-length / distance
// codes; longer codes fall back to the canonical count / sorted-symbol walk.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define NP2_HD __host__ __device__ __forceinline__
#define NP2_HDN inline __host__ __device__ __noinline__
#else
#define NP2_HD inline
#define NP2_HDN inline
#endif

namespace np2 {
namespace infl {

constexpr uint32_t kLitBits = 9, kDistBits = 7;
constexpr uint32_t kInlineMatch = 8;     // matches up to this length are copied by the decoding lane itself (8: see infl_step)
constexpr uint32_t kMaxMember = 65536;  // a BGZF member inflates to at most 64 KiB

struct Tabs {
    // first-level tables, 0 = code longer than the table's bits (or unused); entries carry what the symbol MEANS, so the
    // hot loop does no arithmetic on symbol numbers:
    //   lit_fast : literal b      -> b << 4 | codelen                                   (< 0x1000)
    //              end of block   -> 0x1000 | codelen
    //              length code    -> 0x8000 | extra bits << 12 | (base length - 3) << 4 | codelen
    //              286, 287       -> 0x2000 | codelen (invalid)
    //   dist_fast: distance code  -> base distance << 8 | extra bits << 4 | codelen;  30, 31 -> base 0 (invalid)
    // (the code-length code of a dynamic header is decoded through lit_fast as plain sym << 4 | codelen)
    uint16_t lit_fast[1u << kLitBits];
    uint32_t dist_fast[1u << kDistBits];
    uint16_t lit_sym[288];                // symbols ordered by (code length, symbol)
    uint16_t dist_sym[32];
    uint16_t lit_cnt[16];                 // codes per length
    uint16_t dist_cnt[16];
    uint8_t lens[320];                    // code lengths of the block being set up
};

enum Event : uint32_t { EV_DONE = 0, EV_MATCH = 1, EV_STORED = 2, EV_ERROR = 3 };

struct State {
    // input: the member's raw DEFLATE payload, read as aligned 32-bit words (up to 3 bytes before and 8 bytes after the
    // payload are touched, never interpreted: callers pad their buffers)
    const uint32_t *w;
    uint32_t wi;         // next word
    uint32_t wmax;       // last word index that may be read (the one behind the payload's last word)
    uint32_t ahead;      // w[min(wi, wmax)], loaded one refill early so that its latency hides behind the decoding
    uint64_t buf;        // bit buffer, LSB first
    uint32_t cnt;        // valid bits in buf
    int64_t base_bits;   // bit position of w[0]'s bit 0 relative to the payload's first bit (<= 0)
    uint64_t limit_bits; // payload length in bits
    const uint8_t *payload;
    // output
    uint8_t *out;
    uint32_t pos, cap;
    // block state
    uint32_t inline_max; // matches up to this length (<= kInlineMatch) are copied by infl_step itself
    uint32_t in_block;   // 0 = a block header comes next, 1 = inside a Huffman block
    uint32_t last;       // BFINAL of the current block
};

NP2_HD uint32_t load_word(const uint32_t *p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}
NP2_HD void seat(State &s, uint64_t byte_off) {  // (re)start the bit reader at a byte offset of the payload
    const uintptr_t a = (uintptr_t)(s.payload + byte_off);
    const uint32_t skip = (uint32_t)(a & 3) * 8;
    s.w = reinterpret_cast<const uint32_t *>(a - (a & 3));
    s.base_bits = (int64_t)byte_off * 8 - skip;
    const uint64_t remaining = (s.limit_bits >> 3) - byte_off;
    s.wmax = (uint32_t)(((a & 3) + remaining + 3) >> 2);
    s.buf = (uint64_t)(load_word(s.w) >> skip);
    s.cnt = 32 - skip;
    s.wi = 1;
    s.ahead = load_word(s.w + (1 < s.wmax ? 1 : s.wmax));
}
NP2_HD void init(State &s, const uint8_t *payload, uint32_t clen, uint8_t *out, uint32_t cap,
                 uint32_t inline_max = kInlineMatch) {
    s.payload = payload;
    s.inline_max = inline_max < kInlineMatch ? inline_max : kInlineMatch;
    s.limit_bits = (uint64_t)clen * 8;
    s.out = out;
    s.pos = 0;
    s.cap = cap;
    s.in_block = 0;
    s.last = 0;
    seat(s, 0);
}
// Afterwards at least 33 bits are valid.  A corrupt stream that runs past its payload keeps reading the word behind it:
// the bit count goes on growing, so the end-of-block / end-of-member checks report it, and no read leaves the buffer.
NP2_HD void refill(State &s) {
    if (s.cnt <= 32) {
        s.buf |= (uint64_t)s.ahead << s.cnt;
        s.wi++;
        s.cnt += 32;
        s.ahead = load_word(s.w + (s.wi < s.wmax ? s.wi : s.wmax));
    }
}
NP2_HD uint32_t take(State &s, uint32_t n) {  // n <= 32, n <= cnt
    const uint32_t v = (uint32_t)(s.buf & ((1ull << n) - 1));
    s.buf >>= n;
    s.cnt -= n;
    return v;
}
NP2_HD uint64_t consumed_bits(const State &s) { return (uint64_t)(s.base_bits + 32 * (int64_t)s.wi - (int64_t)s.cnt); }

NP2_HD uint32_t rev_bits(uint32_t v, uint32_t n) {  // reverse the low n bits (n <= 15)
    uint32_t r = 0;
    for (uint32_t i = 0; i < n; i++) r |= ((v >> i) & 1u) << (n - 1 - i);
    return r;
}

enum Kind : uint32_t { K_RAW = 0, K_LIT = 1, K_DIST = 2 };
NP2_HD uint32_t lit_entry(uint32_t sym, uint32_t l) {  // RFC 1951 3.2.5, lengths
    if (sym < 256) return sym << 4 | l;
    if (sym == 256) return 0x1000u | l;
    if (sym > 285) return 0x2000u | l;
    const uint32_t i = sym - 257;
    uint32_t extra = 0, base = 3 + i;
    if (i == 28) {
        base = 258;
    } else if (i >= 8) {
        extra = (i - 4) >> 2;
        base = 3 + ((4 + (i & 3)) << extra);
    }
    return 0x8000u | extra << 12 | (base - 3) << 4 | l;
}
NP2_HD uint32_t dist_entry(uint32_t d, uint32_t l) {  // RFC 1951 3.2.5, distances
    if (d > 29) return l;
    uint32_t extra = 0, base = d + 1;
    if (d >= 4) {
        extra = (d >> 1) - 1;
        base = 1 + ((2 + (d & 1)) << extra);
    }
    return base << 8 | extra << 4 | l;
}
// Canonical Huffman code from n code lengths (RFC 1951 3.2.2).  false = a set of lengths zlib refuses as well: over-
// subscribed, or incomplete — except no code at all, or (not for the code-length code) one single 1-bit code.
// kind: what the first-level entries hold (Tabs); K_RAW and K_LIT fill fast16, K_DIST fast32.
NP2_HDN bool build(const uint8_t *lens, uint32_t n, uint16_t *cnt, uint16_t *sym, uint16_t *fast16, uint32_t *fast32,
                   uint32_t fast_bits, uint32_t kind) {
    for (uint32_t l = 0; l < 16; l++) cnt[l] = 0;
    for (uint32_t i = 0; i < n; i++) cnt[lens[i]]++;
    cnt[0] = 0;
    int32_t left = 1;
    uint16_t offs[16], next_code[16];
    uint32_t code = 0;
    offs[0] = offs[1] = 0;
    next_code[0] = 0;
    for (uint32_t l = 1; l < 16; l++) {
        left = (left << 1) - (int32_t)cnt[l];
        if (left < 0) return false;
        next_code[l] = (uint16_t)code;
        code = (code + cnt[l]) << 1;
        if (l < 15) offs[l + 1] = (uint16_t)(offs[l] + cnt[l]);
    }
    uint32_t max_len = 15;
    while (max_len && !cnt[max_len]) max_len--;
    if (left > 0 && max_len != 0 && (kind == K_RAW || max_len != 1)) return false;
    for (uint32_t i = 0; i < (1u << fast_bits); i++) {
        if (kind == K_DIST) fast32[i] = 0;
        else fast16[i] = 0;
    }
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t l = lens[i];
        if (!l) continue;
        sym[offs[l]++] = (uint16_t)i;
        const uint32_t c = next_code[l]++;
        if (l <= fast_bits) {
            const uint32_t e = kind == K_RAW ? (i << 4 | l) : (kind == K_LIT ? lit_entry(i, l) : dist_entry(i, l));
            for (uint32_t j = rev_bits(c, l); j < (1u << fast_bits); j += 1u << l) {
                if (kind == K_DIST) fast32[j] = e;
                else fast16[j] = (uint16_t)e;
            }
        }
    }
    return true;
}

// A code longer than the first-level table: the canonical walk, one bit at a time (codes of one length are consecutive
// values; the first code of the next length follows the last of this one, shifted).  -> sym << 4 | length, or 0xFFFF0.
NP2_HDN uint32_t decode_long(uint64_t bits, const uint16_t *cnt, const uint16_t *sym) {
    uint32_t code = 0, first = 0, index = 0;
    for (uint32_t l = 1; l < 16; l++) {
        code |= (uint32_t)(bits & 1);
        bits >>= 1;
        const uint32_t c = cnt[l];
        if (code < first + c) return (uint32_t)sym[index + (code - first)] << 4 | l;
        index += c;
        first = (first + c) << 1;
        code <<= 1;
    }
    return 0xFFFF0u;  // no such code
}
// One symbol of the code-length code (raw entries in the first kDistBits-wide slots of lit_fast; its codes are at most
// 7 bits long, so the first level always answers).  Returns the symbol or 0xFFFF (no such code).
NP2_HD uint32_t decode_raw(State &s, const uint16_t *fast, uint32_t fast_bits) {
    const uint32_t e = fast[(uint32_t)s.buf & ((1u << fast_bits) - 1)];
    if (!e) return 0xFFFFu;
    s.buf >>= (e & 15);
    s.cnt -= (e & 15);
    return e >> 4;
}

// The header of a dynamic block (RFC 1951 3.2.7) -> both tables.
NP2_HD bool dynamic_tables(State &s, Tabs &t) {
    refill(s);
    const uint32_t hlit = take(s, 5) + 257, hdist = take(s, 5) + 1, hclen = take(s, 4) + 4;
    if (hlit > 286 || hdist > 30) return false;
    uint8_t cl[19];
    for (uint32_t i = 0; i < 19; i++) cl[i] = 0;
    for (uint32_t i = 0; i < hclen; i++) {
        // the order in which the code-length code's lengths are sent
        // (16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15)
        const uint32_t ord = i < 3 ? 16 + i : (i == 3 ? 0 : ((i & 1) ? (19 - i) >> 1 : 6 + (i >> 1)));
        refill(s);
        cl[ord] = (uint8_t)take(s, 3);
    }
    // the code-length code is decoded through the literal table's slots (they are rebuilt right after)
    if (!build(cl, 19, t.dist_cnt, t.dist_sym, t.lit_fast, nullptr, kDistBits, K_RAW)) return false;
    uint32_t n = 0, prev = 0;
    while (n < hlit + hdist) {
        refill(s);
        const uint32_t c = decode_raw(s, t.lit_fast, kDistBits);
        if (c < 16) {
            t.lens[n++] = (uint8_t)c;
            prev = c;
            continue;
        }
        uint32_t rep, v = 0;
        if (c == 16) {
            if (!n) return false;
            v = prev;
            rep = 3 + take(s, 2);
        } else if (c == 17) {
            rep = 3 + take(s, 3);
        } else if (c == 18) {
            rep = 11 + take(s, 7);
        } else {
            return false;
        }
        if (n + rep > hlit + hdist) return false;
        while (rep--) t.lens[n++] = (uint8_t)v;
        prev = v;
    }
    if (t.lens[256] == 0) return false;  // no end-of-block code
    if (!build(t.lens, hlit, t.lit_cnt, t.lit_sym, t.lit_fast, nullptr, kLitBits, K_LIT)) return false;
    return build(t.lens + hlit, hdist, t.dist_cnt, t.dist_sym, nullptr, t.dist_fast, kDistBits, K_DIST);
}
NP2_HDN bool fixed_tables(Tabs &t) {  // RFC 1951 3.2.6
    for (uint32_t i = 0; i < 288; i++) t.lens[i] = (uint8_t)(i < 144 ? 8 : (i < 256 ? 9 : (i < 280 ? 7 : 8)));
    if (!build(t.lens, 288, t.lit_cnt, t.lit_sym, t.lit_fast, nullptr, kLitBits, K_LIT)) return false;
    for (uint32_t i = 0; i < 30; i++) t.lens[i] = 5;
    // zlib's fixed distance table has all 32 five-bit codes (a complete set); 30 and 31 are rejected when they are used
    t.lens[30] = t.lens[31] = 5;
    return build(t.lens, 32, t.dist_cnt, t.dist_sym, nullptr, t.dist_fast, kDistBits, K_DIST);
}

// Runs until bytes have to be copied or the member ends.
//   EV_MATCH : a = length (kInlineMatch + 1 .. 258), b = distance (1..pos): copy inside the output, then `s.pos += a` and call again
//   EV_STORED: a = length, b = byte offset inside the payload: copy payload -> output, then `s.pos += a` and call again
//   EV_DONE  : the final block ended; s.pos bytes were produced
//   EV_ERROR : not a valid DEFLATE stream for this output size
NP2_HD uint32_t infl_step(State &s, Tabs &t, uint32_t &a, uint32_t &b) {
    for (;;) {
        if (!s.in_block) {
            if (s.last) return consumed_bits(s) <= s.limit_bits ? EV_DONE : EV_ERROR;
            if (consumed_bits(s) + 3 > s.limit_bits) return EV_ERROR;
            refill(s);
            s.last = take(s, 1);
            const uint32_t type = take(s, 2);
            if (type == 0) {
                take(s, s.cnt & 7);  // to the byte boundary
                refill(s);
                const uint32_t len = take(s, 16), nlen = take(s, 16);
                if ((len ^ nlen) != 0xFFFFu) return EV_ERROR;
                const uint64_t at = consumed_bits(s) >> 3;
                if ((at + len) * 8 > s.limit_bits || s.pos + len > s.cap) return EV_ERROR;
                seat(s, at + len);
                if (len) {
                    a = len;
                    b = (uint32_t)at;
                    return EV_STORED;
                }
                continue;
            }
            if (type == 3) return EV_ERROR;
            if (!(type == 1 ? fixed_tables(t) : dynamic_tables(s, t))) return EV_ERROR;
            s.in_block = 1;
        }
        // symbols of the current block
        for (;;) {
            refill(s);
            uint32_t e = t.lit_fast[(uint32_t)s.buf & ((1u << kLitBits) - 1)];
            // e - 1 < 4095: a literal with a first-level code.  One refill (>= 33 bits) covers three of them.
            if (e - 1u < 4095u && s.pos + 3 <= s.cap) {
                uint8_t *o = s.out + s.pos;
                o[0] = (uint8_t)(e >> 4);
                s.buf >>= (e & 15);
                s.cnt -= (e & 15);
                e = t.lit_fast[(uint32_t)s.buf & ((1u << kLitBits) - 1)];
                if (e - 1u < 4095u) {
                    o[1] = (uint8_t)(e >> 4);
                    s.buf >>= (e & 15);
                    s.cnt -= (e & 15);
                    e = t.lit_fast[(uint32_t)s.buf & ((1u << kLitBits) - 1)];
                    if (e - 1u < 4095u) {
                        o[2] = (uint8_t)(e >> 4);
                        s.buf >>= (e & 15);
                        s.cnt -= (e & 15);
                        s.pos += 3;
                        continue;
                    }
                    s.pos += 2;
                } else {
                    s.pos += 1;
                }
                refill(s);  // only adds bits above the ones e was looked up with
            }
            if (!e) {  // a code longer than the first level: its symbol, turned into the same kind of entry
                const uint32_t x = decode_long(s.buf, t.lit_cnt, t.lit_sym);
                e = x == 0xFFFF0u ? 0x2000u : lit_entry(x >> 4, x & 15);
            }
            s.buf >>= (e & 15);
            s.cnt -= (e & 15);
            if (e < 0x1000u) {  // literal
                if (s.pos >= s.cap) return EV_ERROR;
                s.out[s.pos++] = (uint8_t)(e >> 4);
                continue;
            }
            if (!(e & 0x8000u)) {
                if (e >= 0x2000u) return EV_ERROR;  // 286, 287 and "no such code"
                s.in_block = 0;                     // end of block
                if (consumed_bits(s) > s.limit_bits) return EV_ERROR;
                break;
            }
            const uint32_t len = 3 + ((e >> 4) & 0xFFu) + take(s, (e >> 12) & 7u);
            refill(s);
            uint32_t de = t.dist_fast[(uint32_t)s.buf & ((1u << kDistBits) - 1)];
            if (!de) {
                const uint32_t x = decode_long(s.buf, t.dist_cnt, t.dist_sym);
                de = x == 0xFFFF0u ? 0u : dist_entry(x >> 4, x & 15);
            }
            s.buf >>= (de & 15);
            s.cnt -= (de & 15);
            if (!(de >> 8)) return EV_ERROR;  // 30, 31 and "no such code"
            const uint32_t dist = (de >> 8) + take(s, (de >> 4) & 15u);
            if (dist > s.pos || s.pos + len > s.cap) return EV_ERROR;
            if (len <= s.inline_max) {
                // Short matches are most of what a fast deflate level makes of DNA (any 3 bytes of 4-bit SEQ have
                // occurred in the last 32 KiB): copied right here.  Whenever the source lies a full 8 bytes back and the
                // member still has 8 bytes of room, all 8 are copied without looking at the length — the surplus lands
                // on bytes the following symbols write anyway (a valid member fills exactly its ISIZE), the loads are
                // issued together (their L2 latencies overlap) and nothing is predicated.
                uint8_t *d = s.out + s.pos;
                const uint8_t *src = d - dist;
                if (dist >= kInlineMatch && s.pos + kInlineMatch <= s.cap) {
                    const uint8_t v0 = src[0], v1 = src[1], v2 = src[2], v3 = src[3], v4 = src[4], v5 = src[5], v6 = src[6],
                                  v7 = src[7];
                    d[0] = v0, d[1] = v1, d[2] = v2, d[3] = v3, d[4] = v4, d[5] = v5, d[6] = v6, d[7] = v7;
                } else {
                    for (uint32_t i = 0; i < len; i++) d[i] = src[i];
                }
                s.pos += len;
                continue;
            }
            a = len;
            b = dist;
            return EV_MATCH;
        }
    }
}

// The whole member on one host thread (tests; the device kernel has its own driver loop around infl_step).
inline bool inflate_member_host(const uint8_t *payload, uint32_t clen, uint8_t *out, uint32_t cap, uint32_t *produced) {
    State s;
    Tabs t;
    init(s, payload, clen, out, cap);
    for (;;) {
        uint32_t a = 0, b = 0;
        const uint32_t ev = infl_step(s, t, a, b);
        if (ev == EV_DONE) break;
        if (ev == EV_ERROR) return false;
        if (ev == EV_MATCH) {
            for (uint32_t i = 0; i < a; i++) out[s.pos + i] = out[s.pos + i - b];
        } else {
            for (uint32_t i = 0; i < a; i++) out[s.pos + i] = payload[b + i];
        }
        s.pos += a;
    }
    *produced = s.pos;
    return true;
}

}  // namespace infl
}  // namespace np2
