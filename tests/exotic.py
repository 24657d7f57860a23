"""Hand-rolled alignments with the features the seeded generator does not produce (SURVEY §4.1 layer 3): IUPAC and N
bases in reads, lower-case and N stretches in the contig, H/S clip combinations, insertions/deletions right at the
read ends (removed by trim), long indels, =/X and M mixed, reads without any 8-mer anchor, reads touching both contig
ends, secondary/supplementary/duplicate/unmapped flags, zero-length ops."""
import numpy as np

from nextpolish2_b200 import synth

ALPHA = "ACGT"


def _mut(rng, ref, start, length, feat):
    """returns (pos, cigar list, seq) for a read drawn from ref[start:start+length] with random edits"""
    cig, seq = [], []
    p, end = start, min(len(ref), start + length)

    def push(op, n=1):
        if cig and cig[-1][0] == op:
            cig[-1][1] += n
        else:
            cig.append([op, n])
    first = True
    while p < end:
        u = rng.random()
        base = chr(ref[p]).upper()
        if base not in ALPHA:
            base = ALPHA[rng.integers(4)]
        if u < feat["sub"]:
            seq.append(ALPHA[(ALPHA.index(base) + 1 + rng.integers(3)) % 4])
            push("X" if feat["eqx"] else "M")
            p += 1
        elif u < feat["sub"] + feat["ins"] and not first:
            n = int(rng.integers(1, feat["maxindel"]))
            seq.extend(ALPHA[rng.integers(4)] for _ in range(n))
            push("I", n)
        elif u < feat["sub"] + feat["ins"] + feat["del"] and not first:
            n = int(rng.integers(1, feat["maxindel"]))
            n = min(n, end - p)
            push("D", n)
            p += n
        elif u < feat["sub"] + feat["ins"] + feat["del"] + feat["iupac"]:
            seq.append("NRYMKSWN"[rng.integers(8)])
            push("X" if feat["eqx"] else "M")
            p += 1
        else:
            seq.append(base)
            push("=" if feat["eqx"] else "M")
            p += 1
        first = False
    return start, cig, "".join(seq)


def make(seed=1, L=30_000, n_reads=260):
    rng = np.random.default_rng(seed)
    ref = synth.genome(seed + 100, L).copy()
    ref[2000:2300] |= 0x20                     # lower-case stretch: never matches an upper-case read byte in trim
    ref[9000:9040] = ord("N")                  # N run in the contig
    ref[15000] = ord("R")                      # IUPAC in the contig -> code 4 (treated like a gap)
    ref[15500] = ord("m")                      # code 6
    recs = []
    for i in range(n_reads):
        length = int(rng.integers(1200, 6000))
        start = int(rng.integers(0, L - 1000))
        if i % 37 == 0:
            start = 0
        if i % 41 == 0:
            start = max(0, L - length)
        feat = dict(sub=0.004, ins=0.002, dele=0.002, iupac=0.0005, eqx=bool(i % 3 == 0), maxindel=int(rng.choice([2, 4, 40])))
        feat["del"] = feat.pop("dele")
        if i % 29 == 0:
            feat["sub"] = 0.2                  # no 8-mer anchor at all -> discarded by trim
        pos, cig, seq = _mut(rng, ref, start, length, feat)
        flag, mapq = 0, 60
        k = i % 23
        if k == 1:                              # soft clips on both sides
            a, b = int(rng.integers(1, 300)), int(rng.integers(1, 300))
            cig = [["S", a]] + cig + [["S", b]]
            seq = "".join(ALPHA[x] for x in rng.integers(0, 4, a)) + seq + "".join(ALPHA[x] for x in rng.integers(0, 4, b))
        elif k == 2:                            # hard clip then soft clip (is_first handling, main.rs:395-402,429)
            a = int(rng.integers(1, 200))
            cig = [["H", 50], ["S", a]] + cig + [["H", 7]]
            seq = "".join(ALPHA[x] for x in rng.integers(0, 4, a)) + seq
        elif k == 3:                            # insertion / deletion at the very ends (trim removes them)
            cig = [["I", 3]] + cig + [["I", 2]]
            seq = "ACG" + seq + "TT"
        elif k == 4:
            flag = 0x100
        elif k == 5:
            flag = 0x800
        elif k == 6:
            flag = 0x400
        elif k == 7:
            flag = 0x4
        elif k == 8:
            mapq = int(rng.integers(0, 3))
        elif k == 9:
            cig = cig[:1] + [["I", 0]] + cig[1:]  # zero-length op
        elif k == 10:
            flag = 0x10
        recs.append((pos, cig, seq, flag, mapq, "e%d" % i))
    recs.sort(key=lambda r: r[0])
    blob = np.concatenate([synth.bam_record(0, p, [(o, n) for o, n in c], s, flag=f, mapq=q, name=nm)
                           for p, c, s, f, q, nm in recs])
    return ref, blob
