"""A SECOND, independent restatement of the reference's core (ingest -> AlignSeq -> Msa -> DP -> backtrack -> LQ regions
-> candidates -> k-mer scores -> heterozygous regions),
written in plain Python straight from src/main.rs with the reference's own data structures (strings, lists of 3-mers).

The Rust binary cannot be built in this image, so nothing pins the C++ oracle (oracle/np2_oracle.cpp) against the
reference itself.  This file narrows that gap: two restatements written independently from the same source, in
different languages and with different data structures, must agree bit for bit at every stage on small inputs
(tests/test_py_restatement.py).  Slow by design (pure-Python loops): small cases only.

Follows: record filter main.rs:1758-1771; fill_with_cigar 386-440; trim 447-513; AlignSeq::new 279-312;
get_align_tag 314-338; post-trim filter 1796-1813; filter_alignseqs_by_clip 531-574; update_msas 576-589; Kmer 84-184;
Msa 193-241; get_cns_from_align_tags 1645-1687; generate_cns_from_best_score_lq 1555-1643;
generate_lqseqs_from_tags_kmer 1422-1521; retrieve_kmer_count 740-778 + kmer.rs:102-125, 255-287; is_valid_snp 780-801;
get_min_count 803-811; fill_order_stat 813-849; mark_hete_lqseqs 916-946.
"""
import struct

import numpy as np

SEQ_NUM = [65, 67, 71, 84, 45, 78, 77] + [4] * 121
for ch, v in (("A", 0), ("C", 1), ("G", 2), ("T", 3), ("U", 3), ("M", 6), ("N", 5)):
    SEQ_NUM[ord(ch)] = v
    SEQ_NUM[ord(ch.lower())] = v
BAM_SEQ = "=ACMGRSVTWYHKDBN"
HEAD = 15
I64_MIN_HALF = -(1 << 62)  # i64::MIN >> 1


def records(bam):
    """[(tid, pos, mapq, flag, cigar [(op, len)], seq str)] from a raw record blob"""
    out, off, b = [], 0, bytes(bam)
    while off < len(b):
        (bs,) = struct.unpack_from("<i", b, off)
        tid, pos, l_name, mapq, _bin, n_cig, flag, l_seq = struct.unpack_from("<iiBBHHHi", b, off + 4)
        p = off + 36 + l_name
        cig = [(c & 15, c >> 4) for c in struct.unpack_from("<%dI" % n_cig, b, p)]
        p += 4 * n_cig
        seq = "".join(BAM_SEQ[(b[p + (i >> 1)] >> (4 if i % 2 == 0 else 0)) & 15] for i in range(l_seq))
        out.append((tid, pos, mapq, flag, cig, seq))
        off += 4 + bs
    return out


class Alignment:
    def __init__(self):
        self.shift = 0
        self.aln_t_s = self.aln_t_e = self.aln_q_s = self.aln_q_e = 0
        self.q = []
        self.t = []

    def fill_with_cigar(self, cigar, tseq, qseq):  # main.rs:386-440
        qs = ts = 0
        first = True
        for op, l in cigar:
            if op == 4:  # S
                qs += l
                if first:
                    self.aln_q_s = qs
                else:
                    self.aln_q_e = qs - l
            elif op in (0, 7, 8):  # M = X
                self.q.extend(qseq[qs:qs + l])
                qs += l
                self.t.extend(tseq[ts:ts + l])
                ts += l
            elif op == 1:  # I
                self.q.extend(qseq[qs:qs + l])
                qs += l
                self.t.extend("-" * l)
            elif op == 2:  # D
                self.q.extend("-" * l)
                self.t.extend(tseq[ts:ts + l])
                ts += l
            elif op == 5:  # H
                pass
            else:
                raise ValueError("Unknown cigar")
            first = False
        if self.aln_q_e == 0:
            self.aln_q_e = qs
        self.aln_t_e = self.aln_t_s + ts

    def aln_len(self):
        return len(self.t) - self.shift

    def trim(self, n):  # main.rs:447-513
        j = 0
        t, q = self.t, self.q
        for i in range(len(t)):
            if t[i] == q[i]:
                j += 1
                self.aln_t_s += 1
                self.aln_q_s += 1
            else:
                if t[i] != "-":
                    self.aln_t_s += 1
                if q[i] != "-":
                    self.aln_q_s += 1
                j = 0
            if j == n:
                self.aln_t_s -= n
                self.aln_q_s -= n
                self.shift = i + 1 - n
                break
        if j == n:
            j = 0
            for i in range(len(t) - 1, -1, -1):
                if t[i] == q[i]:
                    j += 1
                    self.aln_t_e -= 1
                    self.aln_q_e -= 1
                else:
                    if t[i] != "-":
                        self.aln_t_e -= 1
                    if q[i] != "-":
                        self.aln_q_e -= 1
                    j = 0
                if j == n:
                    self.aln_t_e += n
                    self.aln_q_e += n
                    new_len = i + n
                    if new_len < len(t):
                        del t[new_len:]
                        del q[new_len:]
                    break
        else:
            self.shift = len(t)


class AlignSeq:
    def __init__(self, aln):  # main.rs:279-312
        self.aln_t_s = aln.aln_t_s
        self.aln_t_e = aln.aln_t_s
        self.label = False
        n = (aln.aln_len() + 1) >> 1
        self.align_bases = [0] * (n + 1)
        i = 0
        for tb, qb in zip(aln.t[aln.shift:], aln.q[aln.shift:]):
            b = SEQ_NUM[ord(qb)]
            if tb == "-":
                b |= 8
            elif i != 0:
                self.aln_t_e += 1
            if i & 1 == 0:
                b <<= 4
            self.align_bases[i >> 1] |= b
            i += 1
        self.align_bases[i >> 1] |= 255 if i & 1 == 0 else 15
        self.n_cols = i

    def tags(self):  # get_align_tag main.rs:314-338, as a generator of (q_base, delta, t_pos)
        p, delta, t_pos = 0, 0, 0
        while True:
            t = self.align_bases[p >> 1]
            if p & 1 == 0:
                t >>= 4
            if t & 15 == 15:
                return
            if p != 0:
                if t & 8:
                    delta += 1
                else:
                    delta = 0
                    t_pos += 1
            else:
                t_pos, delta = self.aln_t_s, 0
            p += 1
            yield (t & 7, delta, t_pos)


def ingest(tseq, bam, min_read_len=1000, min_map_len=500, min_map_fra=0.5, min_map_qual=1, max_clip_len=100,
           use_supplementary=False, use_secondary=False):
    """-> alignseqs (index 0 = the ref read), rec_idx per alignseq; main.rs:1732-1817"""
    ref = Alignment()
    ref.aln_t_e = ref.aln_q_e = len(tseq)
    ref.q = list(tseq)
    ref.t = list(tseq)
    alignseqs, rec_idx = [AlignSeq(ref)], [-1]
    pre_tid, pre_pos = 0, 0
    for ri, (tid, pos, mapq, flag, cig, seq) in enumerate(records(bam)):
        if not (tid > pre_tid or pos >= pre_pos):
            raise ValueError("Unsorted input file!")
        rlen = sum(l for op, l in cig if op in (0, 1, 4, 7, 8, 5))
        span = sum(l for op, l in cig if op in (0, 2, 3, 7, 8))
        if flag & 4 or not cig or span == 0:
            span = 1  # bam_endpos
        f32 = float(np.float32(rlen) * np.float32(min_map_fra))  # (rlen as f32 * min_map_fra) as i64, main.rs:1767
        if (flag & 0x404 or mapq <= min_map_qual or rlen <= min_read_len or (flag & 0x100 and not use_secondary)
                or (flag & 0x800 and not use_supplementary) or span < max(min_map_len, int(f32))):
            continue
        aln = Alignment()
        aln.aln_t_s = pos
        aln.fill_with_cigar(cig, tseq[pos:], seq)
        is_clip = aln.aln_q_e - aln.aln_q_s + max_clip_len < rlen
        aln.trim(8)
        if aln.aln_len() <= min_map_len:
            continue
        a = AlignSeq(aln)
        if is_clip:
            if len(tseq) < 500_000:
                continue
            a.label = True
        alignseqs.append(a)
        rec_idx.append(ri)
        pre_tid, pre_pos = tid, pos
    # filter_alignseqs_by_clip main.rs:531-574
    ranges, s, e = [], 0, 0
    for a in alignseqs:
        if a.label:
            continue
        x, y = a.aln_t_s + 50, a.aln_t_e - 50
        if s == e:
            s, e = x, y
        elif x > e:
            ranges.append((s, e))
            s, e = x, y
        elif e < y:
            e = y
    if s != e:
        ranges.append((s, e))
    for a in alignseqs:
        if not a.label:
            continue
        a.label = False
        for rs, re_ in ranges:
            if rs <= a.aln_t_s and a.aln_t_e <= re_:
                a.align_bases = []
                break
            elif a.aln_t_e < rs:
                break
    return alignseqs, rec_idx


def kmer_new(b1, b2, b3):  # main.rs:84-102 -> (bases, delta)
    flags = 0
    if b2[2] == b1[2]:
        flags |= 0b0100
    if b2[2] == b3[2]:
        flags |= 0b0001
    return (((flags << 4 | b1[0]) << 4 | b2[0]) << 4 | b3[0], b1[1])


def kmer_bases(bases, delta, p):  # main.rs:105-184 -> three (q_base, delta, t_pos)
    q1, q2, q3 = bases >> 8 & 15, bases >> 4 & 15, bases & 15
    if bases & 0x5000 == 0x5000:
        return (q1, delta, p), (q2, delta + 1, p), (q3, delta + 2, p)
    if bases & 0x1000:
        return (q1, delta, p - 1), (q2, 0, p), (q3, 1, p)
    if bases & 0x4000:
        return (q1, delta, p - 1), (q2, delta + 1, p - 1), (q3, 0, p)
    return (q1, delta, p - 2), (q2, 0, p - 1), (q3, 0, p)


class K:
    __slots__ = ("bases", "delta", "count", "besti", "score")

    def __init__(self, bases, delta):
        self.bases, self.delta, self.count, self.besti, self.score = bases, delta, 1, 0, 0


def build_msas(L, alignseqs):  # update_msas + sort_msas main.rs:576-589, 227-230
    msas = [[] for _ in range(L)]
    for a in alignseqs:
        if not a.align_bases:
            continue
        b1 = (HEAD, 0, (a.aln_t_s - 1) & 0xFFFFFFFF)
        b2 = (HEAD, 1, (a.aln_t_s - 1) & 0xFFFFFFFF)
        for b3 in a.tags():
            bases, delta = kmer_new(b1, b2, b3)
            for k in msas[b3[2]]:
                if k.bases == bases and k.delta == delta:
                    k.count += 1
                    break
            else:
                msas[b3[2]].append(K(bases, delta))
            b1, b2 = b2, b3
    for m in msas:
        m.sort(key=lambda k: kmer_bases(k.bases, k.delta, 0)[2][1])  # stable, like sort_by_cached_key
    return msas


def coverage(msa):  # main.rs:232-241
    c = 0
    for k in msa:
        if kmer_bases(k.bases, k.delta, 0)[2][1] != 0:
            break
        c += k.count
    return c


def dp(msas):  # get_cns_from_align_tags main.rs:1645-1687 -> the global best 3-mer
    best = K(0, 0)
    L = len(msas)
    for p, msa in enumerate(msas):
        cov = coverage(msa)
        for k in msa:
            base1, base2, _ = kmer_bases(k.bases, k.delta, p)
            besti = 0
            if base2[0] == HEAD:
                score = 10 * k.count - 4 * cov
            else:
                score = I64_MIN_HALF
                b23 = base1[0] << 4 | base2[0]
                d23 = 1 if base1[2] == base2[2] else 0
                for pi, pk in enumerate(msas[base2[2]]):  # Msa::get main.rs:209-225
                    if pk.bases & 0xFF != b23 or (pk.bases >> 12 & 1) != d23:
                        continue
                    pb = kmer_bases(pk.bases, pk.delta, base2[2])
                    if pb[1] != base1 or pb[2] != base2:
                        continue
                    if base2[2] >= 3 and pb[0][0] == HEAD:
                        continue
                    s = pk.score + 10 * k.count - 4 * cov
                    if s > score or (s == score and pb[0][0] != 4):
                        score, besti = s, pi
            k.score, k.besti = score, besti
            if p == L - 1 and score >= best.score:
                best = k
    return best


def backtrack(msas, best):  # generate_cns_from_best_score_lq main.rs:1555-1643
    """-> consensus [(pos, base char, flags)] in ascending order, regions [(start, end)] in the reference's order"""
    cns, regions = [], []
    hq_min_qv, lq_min_length = 95, 2
    has_lq, lq_s, lq_e, p = False, None, 0, 0
    k = best
    _, base2, base3 = kmer_bases(k.bases, k.delta, len(msas) - 1)
    while True:
        if base3[0] != 4:
            cov = coverage(msas[base3[2]])
            qv = k.count * 100 // cov
            cns.append((base3[2], chr(SEQ_NUM[base3[0]]), (1 if qv < hq_min_qv else 0) | (2 if cov < 2 else 0)))
            if cov < 2:
                has_lq, lq_s = False, None
            elif qv < hq_min_qv:
                if lq_s is None:
                    lq_s = p
                lq_e = p
                has_lq = True
            elif (has_lq and p - lq_e > 2 * lq_min_length and cns[p - 1][0] != cns[p - 2][0]
                  and cns[p - 1][1] != cns[p - 2][1]):
                lq_e = p - 2
                lq_s = lq_s - lq_min_length if lq_s > lq_min_length else 1
                while lq_s > 1 and (cns[lq_s - 1][0] == cns[lq_s][0] or cns[lq_s - 1][1] == cns[lq_s][1]):
                    lq_s -= 1
                if regions and cns[lq_s][0] >= regions[-1][0]:
                    regions[-1] = (cns[lq_e][0], regions[-1][1])
                else:
                    regions.append((cns[lq_e][0], cns[lq_s][0]))  # (start, end)
                has_lq, lq_s = False, None
            p += 1
        if base2[0] == HEAD:
            break
        k = msas[base2[2]][k.besti]
        _, base2, base3 = kmer_bases(k.bases, k.delta, base2[2])
    cns.reverse()
    return cns, regions


def yak_hash64(key, mask):  # kmer.rs:223-233 (= yak/yak-priv.h:10-21)
    key = (~key + (key << 21)) & mask
    key = key ^ key >> 24
    key = ((key + (key << 3)) + (key << 8)) & mask
    key = key ^ key >> 14
    key = ((key + (key << 2)) + (key << 4)) & mask
    key = key ^ key >> 28
    key = (key + (key << 31)) & mask
    return key


INVALID_KMER = (1 << 64) - 1


def candidates(alignseqs, regions, ksize, max_can=60):  # generate_lqseqs_from_tags_kmer main.rs:1422-1521
    """regions: [(start, end)] in the reference's (descending) order -> per region [(order, seq, kmer hash)]"""
    out = [[] for _ in regions]
    shift, mask, m64 = 2 * (ksize - 1), (1 << (2 * ksize)) - 1, (1 << 64) - 1
    s = len(regions) - 1
    for idx, a in enumerate(alignseqs):
        if not a.align_bases:
            continue
        while s > 0 and regions[s][0] < a.aln_t_s:
            s -= 1
        if regions[s][0] < a.aln_t_s or regions[s][1] > a.aln_t_e:
            continue
        j = s
        while j > 0 and regions[j][1] <= a.aln_t_e:
            j -= 1
        if regions[j][1] > a.aln_t_e:
            j += 1
        bases = []
        for b in a.tags():
            bases.append(b)
            if b[2] > regions[j][1] + ksize:
                break
        for r in range(j, s + 1):
            if len(out[r]) >= max_can:
                continue
            start, end = regions[r]
            l, k0, k1, seq = 0, 0, 0, []
            for q, _d, t_pos in bases[start - a.aln_t_s:]:
                if t_pos >= start and q != 4:
                    if t_pos <= end:
                        seq.append(chr(SEQ_NUM[q]))
                    if l < ksize:
                        k0 = ((k0 << 2) | q) & mask
                        k1 = ((k1 >> 2) | ((3 ^ q) << shift)) & m64
                        l += 1
                    if t_pos > end and l >= ksize:
                        break
            kmer = min(k0, k1) if l >= ksize else INVALID_KMER
            if seq:
                out[r].append((idx, "".join(seq), yak_hash64(kmer, mask) if kmer != INVALID_KMER else INVALID_KMER))
    return out


def iter2kmer(seq, ksize):  # kmer.rs:255-287 (k < 32): canonical 2-bit k-mers, the window restarts at a non-ACGT base
    shift, mask, m64 = 2 * (ksize - 1), (1 << (2 * ksize)) - 1, (1 << 64) - 1
    k0 = k1 = l = 0
    for ch in seq:
        c = SEQ_NUM[ord(ch)]
        if c < 4:
            k0 = ((k0 << 2) | c) & mask
            k1 = ((k1 >> 2) | ((3 ^ c) << shift)) & m64
            l += 1
        else:
            l = 0
        if l >= ksize:
            yield min(k0, k1)


def kscores(cand, table, ksize, min_kmer_count=5):  # retrieve_kmer_count main.rs:740-778
    """cand: per region [(order, seq, kmer hash)]; table: {hash >> 10: count} -> per region [kscore]"""
    mask = (1 << (2 * ksize)) - 1

    def get(h):  # KmerInfo::get after retrieve_kmers: the stored count if it reaches min_count, else 0
        c = table.get(h >> 10, 0)
        return c if c >= min_kmer_count else 0
    out = []
    for region in cand:
        ks = []
        for _order, seq, kmer in region:
            if len(seq) > ksize:
                vals = [get(yak_hash64(x, mask)) for x in iter2kmer(seq, ksize)]
                ks.append(min(vals) if vals else 0)
            elif kmer != INVALID_KMER:
                ks.append(get(kmer))
            else:
                ks.append(0)
        out.append(ks)
    return out


def is_valid_snp(s1, s2):  # main.rs:780-801
    i = j = 0
    while i < len(s1) and j < len(s2):
        if s1[i] != s2[j]:
            return True
        while i + 1 < len(s1) and s1[i] == s1[i + 1]:
            i += 1
        while j + 1 < len(s2) and s2[j] == s2[j + 1]:
            j += 1
        i += 1
        j += 1
    return False


def get_min_count(c):  # main.rs:803-811
    return 3 if c >= 9 else 2 if c >= 6 else 1


def fill_order_stat(region, ks):  # main.rs:813-849 -> stats per candidate, (max1_c, max1_p, max2_c, max2_p)
    n = len(region)
    stats = [0] * n
    max1_c = max1_p = max2_c = max2_p = 0
    for p1 in range(n):
        if ks[p1] <= 0 or stats[p1] > 0:
            continue
        seq = region[p1][1]
        same = [p for p in range(p1, n) if region[p][1] == seq]
        c = len(same)
        for p in same:
            stats[p] = c
        if c > max1_c or (c == max1_c and region[p1][0] == 0):
            max2_c, max2_p = max1_c, max1_p
            max1_c, max1_p = c, p1
        elif max1_p == max2_p or c > max2_c:
            max2_c, max2_p = c, p1
    return stats, (max1_c, max1_p, max2_c, max2_p)


def mark_hete(cand, ks):  # mark_hete_lqseqs main.rs:916-946 -> per region HETE flag; ks is updated in place
    flags = []
    for region, k in zip(cand, ks):
        stats, (max1_c, max1_p, max2_c, max2_p) = fill_order_stat(region, k)
        min_c = get_min_count(len(region))
        hete = (max2_c >= min_c
                and (len(region[max1_p][1]) == len(region[max2_p][1]) or (len(region) >= 6 and max2_c >= max1_c // 2))
                and is_valid_snp(region[max1_p][1], region[max2_p][1]))
        if hete:
            for p in range(len(region)):
                if k[p] > 0 and stats[p] < min_c:
                    k[p] = 0
        flags.append(hete)
    return flags
