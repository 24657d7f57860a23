"""A SECOND, independent restatement of the reference's whole per-contig path (ingest -> AlignSeq -> Msa -> DP ->
backtrack -> LQ regions -> candidates -> k-mer scores -> heterozygous regions -> read phasing with Louvain -> seed
choice -> consensus patching -> k-mer re-check -> iteration loop; `polish()` at the end runs all of it),
written in plain Python straight from src/main.rs with the reference's own data structures (strings, lists of 3-mers).

The Rust binary cannot be built in this image, so nothing pins the C++ oracle (oracle/np2_oracle.cpp) against the
reference itself.  This file narrows that gap: two restatements written independently from the same source, in
different languages and with different data structures, must agree bit for bit at every stage on small inputs
(tests/test_py_restatement.py).  Slow by design (pure-Python loops): small cases only.

Follows: record filter main.rs:1758-1771; fill_with_cigar 386-440; trim 447-513; AlignSeq::new 279-312;
get_align_tag 314-338; post-trim filter 1796-1813; filter_alignseqs_by_clip 531-574; update_msas 576-589; Kmer 84-184;
Msa 193-241; get_cns_from_align_tags 1645-1687; generate_cns_from_best_score_lq 1555-1643;
generate_lqseqs_from_tags_kmer 1422-1521; retrieve_kmer_count 740-778 + kmer.rs:102-125, 255-287; is_valid_snp 780-801;
get_min_count 803-811; fill_order_stat 813-849; mark_hete_lqseqs 916-946; fill_seed_lqseqs 862-914;
phase_reads_by_lqseqs 948-1015; update_consensus_with_lqseqs 1017-1058; reupdate_consensus_with_lqseqs 1060-1420;
the iteration loop 1821-1838; utils/louvain.rs.
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


def yak_hash64_64(key):  # kmer.rs:235-244
    m = (1 << 64) - 1
    key = (~key + (key << 21)) & m
    key = key ^ key >> 24
    key = ((key + (key << 3)) + (key << 8)) & m
    key = key ^ key >> 14
    key = ((key + (key << 2)) + (key << 4)) & m
    key = key ^ key >> 28
    key = (key + (key << 31)) & m
    return key


def to_hash(kmer, ksize):  # KmerInfo::to_hash kmer.rs:102-110: long k-mers come out of iter2kmer already hashed
    return yak_hash64(kmer, (1 << (2 * ksize)) - 1) if ksize < 32 else kmer


def iter2kmer(seq, ksize):  # kmer.rs:255-310: canonical k-mers, the window restarts at a non-ACGT base
    if ksize >= 32:  # two bit planes per strand, kmer.rs:288-309; yak_hash_long 246-249
        shift, mask = ksize - 1, (1 << ksize) - 1
        x, l = [0, 0, 0, 0], 0
        for ch in seq:
            c = SEQ_NUM[ord(ch)]
            if c < 4:
                x[0] = (x[0] << 1 | (c & 1)) & mask
                x[1] = (x[1] << 1 | (c >> 1)) & mask
                x[2] = x[2] >> 1 | (1 - (c & 1)) << shift
                x[3] = x[3] >> 1 | (1 - (c >> 1)) << shift
                l += 1
            else:
                l = 0
                x = [0, 0, 0, 0]
            if l >= ksize:
                j = 0 if x[1] < x[3] else 1
                yield (yak_hash64_64(x[j << 1]) + yak_hash64_64(x[j << 1 | 1])) & ((1 << 64) - 1)
        return
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


# ------------------------------------------------------------------------------------------------------------------
# The rest of the per-contig path: seed choice, read phasing (Louvain), consensus patching, the k-mer re-check and
# the iteration loop.  Follows main.rs:851-914 (fill_seed_lqseqs), 707-719 (retain_sort_seqs), 948-1015
# (phase_reads_by_lqseqs), 1017-1058 (update_consensus_with_lqseqs), 1060-1420 (reupdate_consensus_with_lqseqs),
# 1524-1553 (the tail of generate_lqseqs_from_tags_kmer), 1821-1838 (the iteration loop) and utils/louvain.rs.
# Where the reference iterates an FxHashMap and the order can reach the result (louvain.rs:123, 145-165, 199) the
# keys are visited in ascending order: the same documented choice as oracle/np2_oracle.cpp.
LABLE_TEMP, LABLE_SUCC, LABLE_HETE, LABLE_RECH = 0x01, 0x80, 0x40, 0x20


class LqSeq:
    __slots__ = ("order", "kscore", "kmer", "seq")

    def __init__(self, order, kscore, kmer, seq):
        self.order, self.kscore, self.kmer, self.seq = order, kscore, kmer, seq


class LqSeqs:
    def __init__(self, start, end, seqs):
        self.lable, self.start, self.end, self.sudoseed, self.seqs = 0, start, end, "", seqs


def order_stats(lq):  # fill_order_stat main.rs:813-849 on LqSeqs, with the order -> count map it fills
    n = len(lq.seqs)
    stats, order_stat = [0] * n, {}
    max1_c = max1_p = max2_c = max2_p = 0
    for p1, s in enumerate(lq.seqs):
        if s.kscore <= 0 or stats[p1] > 0:
            continue
        c = sum(1 for x in lq.seqs[p1:] if x.seq == s.seq)
        order_stat[s.order] = c
        for p2 in range(p1, n):
            if lq.seqs[p2].seq == s.seq:
                stats[p2] = c
        if c > max1_c or (c == max1_c and s.order == 0):
            max2_c, max2_p = max1_c, max1_p
            max1_c, max1_p = c, p1
        elif max1_p == max2_p or c > max2_c:
            max2_c, max2_p = c, p1
    return stats, (max1_c, max1_p, max2_c, max2_p), order_stat


def no_dupseq(lq):  # main.rs:851-860
    for p1 in range(1, len(lq.seqs)):
        for s2 in lq.seqs[p1 + 1:]:
            if lq.seqs[p1].seq == s2.seq:
                return False
    return True


def retain_sort_seqs(lq, stat, min_c):  # main.rs:707-719
    lq.seqs.sort(key=lambda v: -stat.get(v.order, 0))  # stable
    c = 0
    for v in lq.seqs:
        if stat.get(v.order, 0) < min_c:
            break
        c += 1
    del lq.seqs[c:]


def fill_seed_lqseqs(lqseqs, max_indel_len):  # main.rs:862-914
    for lq in lqseqs:
        _, (max1_c, max1_p, _, _), order_stat = order_stats(lq)
        lq.sudoseed = lq.seqs[max1_p].seq
        lq.lable |= LABLE_SUCC | LABLE_RECH
        min_c = get_min_count(len(lq.seqs))
        assert lq.seqs[0].order == 0, "the first lqseq is not ref."
        if 0 in order_stat:
            if 1 < order_stat[0] < min_c:
                order_stat[0] = min_c
        elif sum(1 for x in lq.seqs if x.seq == lq.seqs[0].seq) > 1:
            order_stat[0] = min_c
        if max1_p != 0 and max1_c < min_c and (max1_c > 1 or no_dupseq(lq)):
            assert lq.seqs[max1_p].order in order_stat  # .unwrap()
            order_stat[lq.seqs[max1_p].order] = min_c
            order_stat[0] = min_c
        elif max1_c < min_c:
            order_stat[0] = min_c
        retain_sort_seqs(lq, order_stat, min_c)
        skip_long = abs(len(lq.sudoseed) - len(lq.seqs[0].seq)) > max_indel_len
        if len(lq.seqs) <= 1 or skip_long:
            if lq.seqs or skip_long:
                lq.sudoseed = lq.seqs[0].seq
            lq.lable ^= LABLE_RECH
            lq.seqs = []


def mark_hete_lqseqs(lqseqs):  # main.rs:916-946
    for lq in lqseqs:
        stats, (max1_c, max1_p, max2_c, max2_p), _ = order_stats(lq)
        min_c = get_min_count(len(lq.seqs))
        if (max2_c >= min_c
                and (len(lq.seqs[max1_p].seq) == len(lq.seqs[max2_p].seq) or (len(lq.seqs) >= 6 and max2_c >= max1_c // 2))
                and is_valid_snp(lq.seqs[max1_p].seq, lq.seqs[max2_p].seq)):
            lq.lable |= LABLE_HETE
            for p, s in enumerate(lq.seqs):
                if s.kscore > 0 and stats[p] < min_c:
                    s.kscore = 0


# ---- utils/louvain.rs
def insert_data(data, k1, k2, v):  # louvain.rs:272-279
    d = data.setdefault(k1, {})
    d[k2] = d.get(k2, 0.0) + v


def assign_data(data, k1, k2, v):  # louvain.rs:281-288
    data.setdefault(k1, {})[k2] = v


class LvNode:
    def __init__(self, id_, weight, nodes):
        self.id, self.weight, self.nodes = id_, weight, set(nodes)


class Louvain:
    def __init__(self, data):  # louvain.rs:60-70
        self.data = data
        self.communities = {v: {v} for v in data}
        self.node = {v: LvNode(v, 0.0, [v]) for v in data}

    def first_stage(self):  # louvain.rs:72-117
        mod_inc = False
        visit = sorted(self.data)
        while True:
            can_stop = True
            for v in visit:
                v_nid = self.node[v].id
                node_ids = {}
                for w in self.data[v]:
                    w_nid = self.node[w].id
                    if w_nid in node_ids:
                        continue
                    com = self.communities[w_nid]
                    node_ids[w_nid] = sum(x for k, x in self.data[v].items() if k in com)
                if node_ids:
                    # max_by(weight, then the smaller id is the greater): the last maximum wins among exact equals,
                    # and there are none because the ids differ
                    best = max(node_ids.items(), key=lambda kv: (kv[1], -kv[0]))
                    if best[1] > 0.0 and best[0] != v_nid:
                        self.node[v].id = best[0]
                        self.communities[best[0]].add(v)
                        self.communities[v_nid].discard(v)
                        can_stop = False
                        mod_inc = True
            if can_stop:
                return mod_inc

    def second_stage(self):  # louvain.rs:119-195
        node, communities, decluster = {}, {}, []
        for cid in sorted(self.communities):
            nodes = self.communities[cid]
            if not nodes:
                continue
            nn = LvNode(cid, 0.0, [])
            for nid in nodes:
                vertex = self.node[nid]
                nn.nodes |= vertex.nodes
                nn.weight += vertex.weight
                for k, v in self.data.get(nid, {}).items():
                    if k in nodes:
                        nn.weight += v / 2.0
            if nn.weight < 0.0:
                decluster.append(cid)
            else:
                communities[cid] = {cid}
                node[cid] = nn
        for cid in decluster:
            nodes = self.communities.pop(cid)
            for nid in sorted(nodes):
                new_nid = nid
                while new_nid in communities or new_nid in node:
                    new_nid += 1
                communities[new_nid] = {new_nid}
                node[new_nid] = LvNode(new_nid, self.node[nid].weight, self.node[nid].nodes)
                self.communities[new_nid] = {nid}
        data = {}
        live = [(k, v) for k, v in sorted(self.communities.items()) if v]
        for nid1, nodes1 in live:
            for nid2, nodes2 in live:
                if nid2 <= nid1:
                    continue
                w = 0.0
                for vid in nodes1:
                    for k, v in self.data.get(vid, {}).items():
                        if k in nodes2:
                            w += v
                if w != 0.0:
                    insert_data(data, nid1, nid2, w)
                    insert_data(data, nid2, nid1, w)
        nxt = Louvain.__new__(Louvain)
        nxt.data, nxt.communities, nxt.node = data, communities, node
        return nxt

    def get_communities(self):  # louvain.rs:197-245
        out = []
        for cid in sorted(self.communities):
            nodes = self.communities[cid]
            if not nodes:
                continue
            weight, new_nodes = 0.0, set()
            for vid in nodes:
                v = self.node[vid]
                new_nodes |= v.nodes
                weight += v.weight
                for k, x in self.data.get(vid, {}).items():
                    if k in nodes:
                        weight += x / 2.0
            out.append(LvNode(cid, weight, new_nodes))
        data = {}
        for c1 in out:
            for c2 in out:
                if c2.id <= c1.id:
                    continue
                w = 0.0
                for n1 in self.communities[c1.id]:
                    for n2 in self.communities[c2.id]:
                        w += self.data.get(n1, {}).get(n2, 0.0)
                if w != 0.0:
                    assert w < 0.0, "the weight of two conflicting community is not less than 0"
                    insert_data(data, c1.id, c2.id, w)
                    insert_data(data, c2.id, c1.id, w)
        return data, out

    def execute(self):  # louvain.rs:247-256
        lv = self
        while lv.first_stage():
            lv = lv.second_stage()
        return lv.get_communities()


def phase_communities(data, ref_weight):  # louvain.rs:290-356
    cdata, communities = Louvain(data).execute()
    if ref_weight is not None:
        def stat(nodes):
            count, weight = 0, 0.0
            for n in nodes:
                v = ref_weight.get(n)
                if v is not None:
                    count += 1 if v > 0 else -1 if v < 0 else 0
                    weight += v
            return (count, weight)
        keyed = [(stat(c.nodes), c) for c in communities]
        keyed.sort(key=lambda kc: (-kc[0][0], -kc[0][1]))  # Reverse((count, weight)), stable
        communities = [c for _, c in keyed]
    else:
        communities.sort(key=lambda c: -c.weight)
    invalid = set()
    for p, c in enumerate(communities):
        if c.id in invalid:
            continue
        vs = cdata.get(c.id)
        if vs is not None:
            for chk in communities[p + 1:]:
                if chk.id not in invalid and chk.id in vs:
                    invalid.add(chk.id)
    out = []
    for c in communities:
        if c.id in invalid:
            out.extend(c.nodes)
    return out


def phase_reads_by_lqseqs(lqseqs, asref, use_all_reads):  # main.rs:948-1015
    data, dif, ref_data, invalid_ids = {}, {}, {}, set()
    for lq in lqseqs:
        if not lq.lable & LABLE_HETE:
            continue
        for i, s1 in enumerate(lq.seqs):
            if s1.kscore == 0:
                continue
            for s2 in lq.seqs[i + 1:]:
                if s2.kscore == 0:
                    continue
                w = 1.0 if s1.seq == s2.seq else -1.0
                if s1.order == 0:
                    if asref:
                        insert_data(ref_data, s1.order, s2.order, w)
                    if w < 0 and not use_all_reads:
                        invalid_ids.add(s2.order)
                    continue
                assert s2.order != 0, "seq2 order is equal to 0"
                if w == -1.0:
                    insert_data(dif, s1.order, s2.order, -1.0)
                    insert_data(dif, s2.order, s1.order, -1.0)
                insert_data(data, s1.order, s2.order, w)
                insert_data(data, s2.order, s1.order, w)
    for n1, vs in dif.items():
        for n2, w in vs.items():
            if w <= -3.0:
                assign_data(data, n1, n2, w)
    if not use_all_reads:
        data = {k: {k2: w for k2, w in vs.items() if k2 not in invalid_ids} for k, vs in data.items()
                if k not in invalid_ids}
    out = phase_communities(data, ref_data[0] if ref_data else None)
    return out + sorted(invalid_ids)


def update_consensus_with_lqseqs(lqseqs, consensus, lable):  # main.rs:1017-1058; consensus = [(pos, base)]
    def next_idx(i):  # get_lqseqs_next_idx_by_lable: usize arithmetic, "below 0" is any index >= len
        i -= 1
        while 0 <= i < len(lqseqs) and not lqseqs[i].lable & lable:
            i -= 1
        return i
    out, i, li = [], 0, next_idx(len(lqseqs))
    while i < len(consensus):
        p = consensus[i][0]
        if 0 <= li < len(lqseqs) and p == lqseqs[li].start:
            out.extend((p, b) for b in lqseqs[li].sudoseed)
            while i < len(consensus) and consensus[i][0] <= lqseqs[li].end:
                i += 1
            li = next_idx(li)
        else:
            out.append(consensus[i])
            i += 1
    return out


def reupdate_consensus_with_lqseqs(lqseqs, consensus, get, ksize, iter_count):  # main.rs:1060-1420
    """get(hash) = KmerInfo::get after retrieve_kmers(min_kmer_count) for this table"""
    idx = [0]

    def pos(i):
        assert 0 <= i < len(consensus), "index out of range (the reference panics)"
        return consensus[i][0]

    def region(s, e):  # iter_consensus_region: the bases strictly between s and e
        i = idx[0]
        while pos(i) <= s:
            i += 1
        while pos(i) > s:
            i -= 1
        i += 1
        si = i
        while pos(i) >= e:
            i -= 1
        while pos(i) < e:
            i += 1
        i -= 1
        idx[0] = i
        return si, i + 1

    def extend(p, l, toleft):  # iter_consensus_extend: l bases left of / right of position p
        i = idx[0]
        if toleft:
            while pos(i) >= p:
                i -= 1
            while pos(i) < p:
                i += 1
            idx[0] = i
            return (i - l if i > l else 0), i
        while pos(i) <= p:
            i += 1
        while pos(i) > p:
            i -= 1
        idx[0] = i
        return i + 1, (i + l + 1 if i + l < len(consensus) else len(consensus))

    def bases(si, ei):
        return "".join(b for _, b in consensus[si:ei])

    def chain(combo, sj, left, right):  # iter_chain_lqseqs
        s = bases(*left)
        for i, (_, seq) in enumerate(combo):
            s += seq
            if i < len(combo) - 1:
                a, b = lqseqs[rech[sj + i]].end, lqseqs[rech[sj + i + 1]].start
                if a + 1 != b:
                    s += bases(*region(a, b))
            else:
                s += bases(*right)
        return s

    def score(s):
        vals = [get(to_hash(x, ksize)) for x in iter2kmer(s, ksize)]
        return min(vals) if vals else 0

    import itertools
    rech = [i for i in range(len(lqseqs) - 1, -1, -1) if lqseqs[i].lable & LABLE_RECH]
    # (the first pass of the reference only collects the k-mers to look up; `get` already answers for any k-mer, but
    # the cursor walk of that pass is repeated because it moves idx, main.rs:1193-1261)
    for collect in (True, False):
        idx[0] = 0
        sj = 0
        while sj < len(rech):
            ej = sj + 1
            while ej < len(rech) and lqseqs[rech[ej]].start < lqseqs[rech[ej - 1]].end + ksize:
                ej += 1
                if ej > sj + 5:
                    break
            left = extend(lqseqs[rech[sj]].start, ksize - 1, True)
            right = extend(lqseqs[rech[ej - 1]].end, ksize - 1, False)
            if ej == sj + 1:
                if not collect:
                    for s in lqseqs[rech[sj]].seqs:
                        s.kscore = score(bases(*left) + s.seq + bases(*right))
            else:
                buf = []
                for combo in itertools.product(*[list(enumerate(x.seq for x in lqseqs[rech[x_]].seqs)) for x_ in range(sj, ej)]):
                    ks = score(chain(combo, sj, left, right))
                    if ks > 0 and not collect:
                        for i, (p, _) in enumerate(combo):
                            buf.append((rech[sj + i], p, ks))
                if not collect:
                    for x_ in range(sj, ej):
                        for s in lqseqs[rech[x_]].seqs:
                            s.kscore = 0
                    for i, p, ks in buf:
                        lqseqs[i].seqs[p].kscore = ks
            sj = ej
    for lq in lqseqs:
        if not lq.lable & LABLE_RECH:
            continue
        c = valid = 0
        for p, s in enumerate(lq.seqs):
            if s.kscore != 0:
                if c == 0 or s.order == 0:
                    c = p + 1
                valid += 1
        if valid > 1:
            lq.lable |= LABLE_TEMP
        if c != 0:
            lq.sudoseed = lq.seqs[c - 1].seq
        elif iter_count == 1:
            i = 0
            for p, s in enumerate(lq.seqs):
                if s.order == 0:
                    i = p
                    break
            lq.sudoseed = lq.seqs[i].seq
    consensus = update_consensus_with_lqseqs(lqseqs, consensus, LABLE_RECH)
    for lq in lqseqs:
        if lq.lable & LABLE_RECH:
            lq.lable ^= LABLE_TEMP if lq.lable & LABLE_TEMP else LABLE_RECH
    return consensus


def polish(tseq, bam, tables, iter_count=2, asref=True, use_all_reads=False, max_indel_len=20, min_kmer_count=5,
           **ingest_kw):
    """One contig through the whole path (main.rs:1727-1838).  tables = [(ksize, {hash >> 10: count})]; the smallest ksize < 32 (main.rs:1432-1434).
    -> (consensus [(pos, base)], [sorted read indices blanked by each non-final iteration])"""
    tables = sorted(tables, key=lambda t: t[0])  # option.rs:238
    k0, tab0 = tables[0]

    def getter(tab):
        def get(h):
            c = tab.get(h >> 10, 0)
            return c if c >= min_kmer_count else 0
        return get
    als, _ = ingest(tseq, bam, **ingest_kw)
    dropped, i = [], 0
    while True:
        msas = build_msas(len(tseq), als)
        cns, regions = backtrack(msas, dp(msas))
        consensus = [(p, b) for p, b, _ in cns]
        final = i + 1 == iter_count
        if regions:
            cand = candidates(als, regions, k0)
            ks = kscores(cand, tab0, k0, min_kmer_count)
            lqseqs = [LqSeqs(s, e, [LqSeq(o, k, h, q) for (o, q, h), k in zip(c, kk)])
                      for (s, e), c, kk in zip(regions, cand, ks)]
            if final:
                fill_seed_lqseqs(lqseqs, max_indel_len)
                consensus = update_consensus_with_lqseqs(lqseqs, consensus, LABLE_SUCC)
                for p, (k, tab) in enumerate(tables):
                    consensus = reupdate_consensus_with_lqseqs(lqseqs, consensus, getter(tab), k, p + 1)
            else:
                mark_hete_lqseqs(lqseqs)
                bad = phase_reads_by_lqseqs(lqseqs, asref, use_all_reads)
                for r in bad:
                    als[r].align_bases = []
                dropped.append(sorted(set(bad)))
        elif not final:
            dropped.append([])
        if final:
            return consensus, dropped
        i += 1


def phase_from_pairs(pairs, asref=True, use_all_reads=False):
    """phase_reads_by_lqseqs (main.rs:948-1015) fed with (a, b, #agree, #differ) per read pair (a < b; a == 0 is the
    ref read) instead of LqSeqs: the per-site +1 / -1 updates are replayed one by one."""
    data, dif, ref_data, invalid_ids = {}, {}, {}, set()
    for a, b, agree, differ in pairs:
        for w in [1.0] * agree + [-1.0] * differ:
            if a == 0:
                if asref:
                    insert_data(ref_data, a, b, w)
                if w < 0 and not use_all_reads:
                    invalid_ids.add(b)
                continue
            if w == -1.0:
                insert_data(dif, a, b, -1.0)
                insert_data(dif, b, a, -1.0)
            insert_data(data, a, b, w)
            insert_data(data, b, a, w)
    for n1, vs in dif.items():
        for n2, w in vs.items():
            if w <= -3.0:
                assign_data(data, n1, n2, w)
    if not use_all_reads:
        data = {k: {k2: w for k2, w in vs.items() if k2 not in invalid_ids} for k, vs in data.items()
                if k not in invalid_ids}
    return sorted(set(phase_communities(data, ref_data[0] if ref_data else None)) | invalid_ids)
