"""BASELINE.json configs[0]: the reference's bundled test set (test/asm.fa.gz + test/hifi.fasta.gz, k21 + k31 tables
written by the reference's own `yak count`), REAL HiFi reads with real error patterns and real heterozygosity
(the 100 kb excerpt is diploid: ~440 HETE regions, ~240 reads dropped by phasing, the 60-candidate cap triggers).

Two fixtures, both made by tests/golden/make_c1.py in the build container:
  c1_40k  committed: first 40 kb of the contig, the reads aligned inside it, tables cut to those reads' k-mers
  _c1     git-ignored (travels with the snapshot): the whole contig, all reads, the unmodified yak files
"""
import gzip
import hashlib
import os

import numpy as np
import pytest

import common
import oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_fixture(name):
    d = os.path.join(GOLD, name)
    if not os.path.isdir(d) or not os.path.exists(os.path.join(d, "k21.yak")):
        pytest.skip("fixture %s not built (tests/golden/make_c1.py needs /root/reference)" % name)
    if name == "c1_40k":
        contig = np.frombuffer(gzip.open(os.path.join(d, "contig.bin.gz")).read(), np.uint8)
        bam = np.frombuffer(gzip.open(os.path.join(d, "records.bin.gz")).read(), np.uint8)
    else:
        contig = np.fromfile(os.path.join(d, "contig.bin"), np.uint8)
        bam = np.fromfile(os.path.join(d, "records.bin"), np.uint8)
    return contig, bam, [os.path.join(d, "k%d.yak" % k) for k in (21, 31)]


def _records(bam):
    off = 0
    while off < len(bam):
        bs = int(np.frombuffer(bam[off:off + 4], "<i4")[0])
        yield off, bs
        off += 4 + bs


def test_c1_records_are_valid_alignments():
    """The mini-aligner's records: CIGAR query length == l_seq, the M columns are >= 97 % identical to the contig."""
    contig, bam, _ = load_fixture("c1_40k")
    dec = np.frombuffer(b"=ACMGRSVTWYHKDBN", np.uint8)
    n, last_pos = 0, -1
    for off, bs in _records(bam):
        pos = int(np.frombuffer(bam[off + 8:off + 12], "<i4")[0])
        assert pos >= last_pos
        last_pos = pos
        l_name = int(bam[off + 12])
        n_cig = int(np.frombuffer(bam[off + 16:off + 18], "<u2")[0])
        l_seq = int(np.frombuffer(bam[off + 20:off + 24], "<i4")[0])
        cig = np.frombuffer(bam[off + 36 + l_name:off + 36 + l_name + 4 * n_cig], "<u4")
        sq = bam[off + 36 + l_name + 4 * n_cig:off + 36 + l_name + 4 * n_cig + (l_seq + 1) // 2]
        seq = dec[np.stack([sq >> 4, sq & 15], 1).reshape(-1)[:l_seq]]
        q, t, match, cols = 0, pos, 0, 0
        for c in cig:
            op, ln = int(c & 15), int(c >> 4)
            if op == 0:
                match += int((seq[q:q + ln] == contig[t:t + ln]).sum())
                cols += ln
                q += ln
                t += ln
            elif op == 1 or op == 4:
                q += ln
            elif op == 2:
                t += ln
            else:
                raise AssertionError("unexpected op %d" % op)
        assert q == l_seq and t <= len(contig)
        assert match >= 0.97 * cols
        n += 1
    assert n == 216


def test_c1_oracle_pinned_digest():
    """Self-pin (the Rust binary cannot be built here): the oracle's FASTA for the committed C1 window."""
    contig, bam, yaks = load_fixture("c1_40k")
    oj = O.Job(contig, bam, [O.Table.load(p) for p in yaks], O.Opts(min_ctg_len=0), dump_iter=0)
    pos, base = oj.consensus()
    reg = oj.regions()
    assert len(reg["start"]) > 500 and (reg["lable"] != 0).sum() > 50  # real data: many LQ regions, real het sites
    assert len(oj.dropped()) > 20                                      # phasing drops the other haplotype's reads
    assert np.diff(oj.candidates()["roff"]).max() == 60                # the cap of main.rs:30,1474 triggers at ~74x
    fa = O.format_fasta("ptg000005l:21113231-21213230", pos, base)
    want = open(os.path.join(GOLD, "c1_40k", "oracle_fasta.sha256")).read().strip()
    assert hashlib.sha256(fa).hexdigest() == want


@pytest.mark.gpu
@pytest.mark.parametrize("fixture", ["c1_40k", "_c1"])
def test_c1_gpu_identical(ctx, fixture):
    """Every stage and the FASTA, CUDA path vs oracle, tables read from the yak files by np2_yak_load."""
    import nextpolish2_b200 as np2
    contig, bam, yaks = load_fixture(fixture)
    oo, go = common.same_opts()
    ot = [O.Table.load(p) for p in yaks]
    gt = [np2.Table.load(ctx, p) for p in yaks]
    for it in (0, 1):
        oj = O.Job(contig, bam, ot, oo, dump_iter=it)
        gj = np2.Job(ctx, contig, bam, gt, go).upload().run(it)
        common.assert_same_dict("reads", oj.reads(), gj.reads(), ["rec_idx", "t_s", "t_e", "blank", "nib_off", "nib"])
        common.assert_same_dict("msa", oj.msa(), gj.msa(), ["off", "bases", "delta", "count", "besti"])
        common.assert_same_dict("dp", oj.dp_consensus(), gj.dp_consensus())
        common.assert_same_dict("regions", oj.regions(), gj.regions(), ["start", "end", "lable"])
        common.assert_same_dict("cand", oj.candidates(), gj.candidates(), ["roff", "order", "seq_off", "seq", "kmer", "kscore"])
        common.assert_same("dropped", oj.dropped(), gj.dropped())
        opos, obase = oj.consensus()
        gpos, gbase = gj.consensus()
        common.assert_same("final.base", obase, gbase)
        common.assert_same("final.pos", opos, gpos)
        gj.destroy()
    name = "ptg000005l:21113231-21213230"
    assert O.format_fasta(name, opos, obase) == np2.format_fasta(name, gpos, gbase)


@pytest.mark.gpu
@pytest.mark.parametrize("optkw", [{"iter_count": 1}, {"use_all_reads": 1}, {"model": 1}, {"min_kmer_count": 20, "iter_count": 3}])
def test_c1_gpu_options(ctx, optkw):
    import nextpolish2_b200 as np2
    contig, bam, yaks = load_fixture("c1_40k")
    oo, go = common.same_opts(**optkw)
    opos, obase = O.Job(contig, bam, [O.Table.load(p) for p in yaks], oo, dump_iter=-1).consensus()
    gpos, gbase = np2.polish_contig(ctx, contig, bam, [np2.Table.load(ctx, p) for p in yaks], go)
    common.assert_same("final.base", obase, gbase)
    common.assert_same("final.pos", opos, gpos)
