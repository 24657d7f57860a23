"""Builds the BASELINE.json configs[0] fixtures (the reference's bundled test set) in the build container:

    python tests/golden/make_c1.py            # both outputs
    python tests/golden/make_c1.py --full     # only the git-ignored full set (what __graft_entry__.build() runs)

Inputs (read here only, never on the GPU box): /root/reference/test/{asm.fa.gz, hifi.fasta.gz, sr.R1/R2.fastq.gz} and
the reference's own `yak count` compiled into oracle/_ref/yak (the commands of /root/reference/test/hh.sh).  The
reference aligns with minimap2, which this image lacks; the alignment records come from nextpolish2_b200/synth's
mini-aligner instead (exact-k-mer anchors + unit-cost DP between anchors), so they are valid alignments of the real
reads but not minimap2's.

Outputs:
  tests/golden/_c1/        (git-ignored, travels to the GPU box)  the whole 100 kb contig, all 575 reads, the
                           unmodified k21.yak / k31.yak written by the reference `yak count`
  tests/golden/c1_40k/     (committed, ~1.5 MB)  the first 40 kb of the contig, the reads aligned inside it, and the
                           two tables cut down to the k-mers of those reads and the contig (same counts), re-written
                           in the yak dump layout
"""
import argparse
import gzip
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from nextpolish2_b200 import synth  # noqa: E402

REF_TEST = "/root/reference/test/"
YAK = os.path.join(ROOT, "oracle", "_ref", "yak")
WINDOW = 40_000


def records(bam):
    """(offset, size, pos, ref_end) of each record in a raw record blob."""
    out, off = [], 0
    while off < len(bam):
        bs = int(np.frombuffer(bam[off:off + 4], "<i4")[0])
        pos = int(np.frombuffer(bam[off + 8:off + 12], "<i4")[0])
        l_name = int(bam[off + 12])
        n_cig = int(np.frombuffer(bam[off + 16:off + 18], "<u2")[0])
        cig = np.frombuffer(bam[off + 36 + l_name:off + 36 + l_name + 4 * n_cig], "<u4")
        span = int(sum(int(c >> 4) for c in cig if (c & 15) in (0, 2, 3, 7, 8)))
        out.append((off, 4 + bs, pos, pos + span))
        off += 4 + bs
    return out


def read_yak(path):
    """-> (k, hashes u64, counts u16) from a yak dump (yak/htab.c:190-211)."""
    raw = np.fromfile(path, np.uint8)
    assert bytes(raw[:4]) == b"YAK\2"
    k, pre, bits = [int(x) for x in np.frombuffer(raw[4:16], "<u4")]
    assert pre == 10 and bits == 10
    off, hs, cs = 16, [], []
    for b in range(1 << pre):
        cap, size = [int(x) for x in np.frombuffer(raw[off:off + 8], "<u4")]
        keys = np.frombuffer(raw[off + 8:off + 8 + 8 * size], "<u8")
        hs.append(((keys >> np.uint64(10)) << np.uint64(pre)) | np.uint64(b))
        cs.append((keys & np.uint64(1023)).astype(np.uint16))
        off += 8 + 8 * size
    return k, np.concatenate(hs), np.concatenate(cs)


def make_full(out):
    os.makedirs(out, exist_ok=True)
    asm = synth.read_fasta(REF_TEST + "asm.fa.gz")
    reads = synth.read_fasta(REF_TEST + "hifi.fasta.gz")
    contig = np.frombuffer(asm[0][1], np.uint8)
    bam, n = synth.align_reads(contig, reads, threads=8)
    assert n == len(reads)
    contig.tofile(os.path.join(out, "contig.bin"))
    bam.tofile(os.path.join(out, "records.bin"))
    open(os.path.join(out, "name.txt"), "w").write(asm[0][0] + "\n")
    for k in (21, 31):
        p = os.path.join(out, "k%d.yak" % k)
        if not os.path.exists(p):
            cmd = "%s count -o %s -k %d <(zcat %ssr.R*.fastq.gz) <(zcat %ssr.R*.fastq.gz)" % (YAK, p, k, REF_TEST, REF_TEST)
            subprocess.check_call(cmd, shell=True, executable="/bin/bash", stderr=subprocess.DEVNULL)
    return contig, bam, reads


def make_window(full_dir, out, contig, bam):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    os.makedirs(out, exist_ok=True)
    sub = contig[:WINDOW]
    keep = [r for r in records(bam) if r[3] <= WINDOW]
    blob = np.concatenate([bam[o:o + s] for o, s, _, _ in keep])
    # QUAL is 0xFF filler and SEQ uses 4 of 16 codes: gzip keeps the committed file small
    with gzip.open(os.path.join(out, "records.bin.gz"), "wb", 9) as f:
        f.write(blob.tobytes())
    with gzip.open(os.path.join(out, "contig.bin.gz"), "wb", 9) as f:
        f.write(sub.tobytes())
    # k-mers of the contig window and of every kept read (either strand: the hashes are canonical)
    seqs = [sub.tobytes()]
    dec = np.frombuffer(b"=ACMGRSVTWYHKDBN", np.uint8)
    for o, s, _, _ in keep:
        l_name = int(bam[o + 12])
        n_cig = int(np.frombuffer(bam[o + 16:o + 18], "<u2")[0])
        l_seq = int(np.frombuffer(bam[o + 20:o + 24], "<i4")[0])
        sq = bam[o + 36 + l_name + 4 * n_cig:o + 36 + l_name + 4 * n_cig + (l_seq + 1) // 2]
        nib = np.stack([sq >> 4, sq & 15], 1).reshape(-1)[:l_seq]
        seqs.append(dec[nib].tobytes())
    for k in (21, 31):
        _, h, c = read_yak(os.path.join(full_dir, "k%d.yak" % k))
        want = np.unique(np.concatenate([O.seq_hashes(s, k) for s in seqs]))
        m = np.isin(h, want)
        synth.write_yak(os.path.join(out, "k%d.yak" % k), k, h[m], c[m])
        print("k%d: %d of %d keys kept" % (k, int(m.sum()), len(h)))
    print("window: %d reads, %d record bytes" % (len(keep), len(blob)))
    # self-pin: the oracle's FASTA for this window (tests/test_c1_bundled.py::test_c1_oracle_pinned_digest)
    import hashlib
    tabs = [O.Table.load(os.path.join(out, "k%d.yak" % k)) for k in (21, 31)]
    pos, base = O.Job(sub, blob, tabs, O.Opts(min_ctg_len=0), dump_iter=-1).consensus()
    name = open(os.path.join(full_dir, "name.txt")).read().strip()
    open(os.path.join(out, "oracle_fasta.sha256"), "w").write(hashlib.sha256(O.format_fasta(name, pos, base)).hexdigest() + "\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    a = ap.parse_args()
    synth.build()
    full = os.path.join(HERE, "_c1")
    contig, bam, _ = make_full(full)
    if not a.full:
        make_window(full, os.path.join(HERE, "c1_40k"), contig, bam)


if __name__ == "__main__":
    main()
