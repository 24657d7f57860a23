"""-S / --use_secondary (src/utils/secondary.rs:8-158, main.rs:1775-1788): secondary alignments have no SEQ in the BAM;
it is recovered from the primary record of the same read (anywhere in the file), in the read's original orientation,
and reverse-complemented again when the secondary record is on the reverse strand.

Construction used here: take a seeded data set, move the record of every third read to another reference (as the
read's PRIMARY alignment, on a randomly chosen strand) and leave a SEQ-less SECONDARY record with the same CIGAR in
its place.  Recovering the sequences must give back exactly the original records' SEQ, so polishing the rewritten blob
with -S must equal polishing the original blob without it.
"""
import os
import subprocess

import numpy as np
import pytest

import common
import oracle as O
import nextpolish2_b200 as np2
from nextpolish2_b200 import synth

COMP = np.arange(16, dtype=np.uint8)
COMP[[1, 8, 2, 4]] = [8, 1, 4, 2]  # A<->T, C<->G on BAM 4-bit codes; everything else unchanged (secondary.rs:72-82)


def records(bam):
    off = 0
    while off < len(bam):
        bs = int(np.frombuffer(bam[off:off + 4], "<i4")[0])
        yield bam[off:off + 4 + bs]
        off += 4 + bs


def fields(rec):
    l_name = int(rec[12])
    n_cig = int(np.frombuffer(rec[16:18], "<u2")[0])
    flag = int(np.frombuffer(rec[18:20], "<u2")[0])
    l_seq = int(np.frombuffer(rec[20:24], "<i4")[0])
    head = 36 + l_name + 4 * n_cig
    return l_name, n_cig, flag, l_seq, head


def unpack(rec):
    _, _, _, l_seq, head = fields(rec)
    sq = rec[head:head + (l_seq + 1) // 2]
    return np.stack([sq >> 4, sq & 15], 1).reshape(-1)[:l_seq]


def build(rec, codes, flag=None, ref_id=None, pos=None):
    """the record with SEQ replaced by `codes` (QUAL 0xFF), optionally a new flag / reference / position"""
    l_name, n_cig, f0, l_seq, head = fields(rec)
    n = len(codes)
    c = np.concatenate([codes, np.zeros(n & 1, np.uint8)]).astype(np.uint8)
    packed = (c[0::2] << 4 | c[1::2]).astype(np.uint8)
    tail = rec[head + (l_seq + 1) // 2 + l_seq:]
    body = np.concatenate([rec[4:head], packed, np.full(n, 0xFF, np.uint8), tail])
    body[16:20] = np.frombuffer(np.array([n], "<i4").tobytes(), np.uint8)
    if flag is not None:
        body[14:16] = np.frombuffer(np.array([flag], "<u2").tobytes(), np.uint8)
    if ref_id is not None:
        body[0:4] = np.frombuffer(np.array([ref_id], "<i4").tobytes(), np.uint8)
    if pos is not None:
        body[4:8] = np.frombuffer(np.array([pos], "<i4").tobytes(), np.uint8)
    return np.concatenate([np.frombuffer(np.array([len(body)], "<i4").tobytes(), np.uint8), body])


def split_secondary(bam, seed=1):
    """-> (contig-0 blob with every third record turned into a SEQ-less secondary, blob of their primaries on ref 1)"""
    rng = np.random.default_rng(seed)
    here, elsewhere = [], []
    for i, rec in enumerate(records(bam)):
        _, _, flag, _, _ = fields(rec)
        if i % 3 != 1 or flag & 0x900:
            here.append(rec)
            continue
        s = unpack(rec)
        orig = COMP[s[::-1]] if flag & 0x10 else s            # the read as sequenced
        g = 0x10 if rng.integers(2) else 0                    # strand of its primary alignment
        elsewhere.append(build(rec, COMP[orig[::-1]] if g else orig, flag=g, ref_id=1, pos=int(rng.integers(1000))))
        here.append(build(rec, np.empty(0, np.uint8), flag=flag | 0x100))
    return np.concatenate(here), np.concatenate(elsewhere)


def expected_fill(blob, primaries):
    """independent restatement of secondary.rs:85-150 + main.rs:1775-1783 on record blobs"""
    ids = set()
    for rec in records(blob):
        l_name, _, flag, _, _ = fields(rec)
        if flag & 0x100:
            ids.add(bytes(rec[36:36 + l_name - 1]))
    seqs = {}
    for src in (blob, primaries):
        for rec in records(src):
            l_name, _, flag, _, _ = fields(rec)
            q = bytes(rec[36:36 + l_name - 1])
            if q in ids and not flag & 0x900:
                s = unpack(rec)
                assert q not in seqs
                seqs[q] = COMP[s[::-1]] if flag & 0x10 else s
    out = []
    for rec in records(blob):
        l_name, _, flag, _, _ = fields(rec)
        if flag & 0x100:
            s = seqs[bytes(rec[36:36 + l_name - 1])]
            rec = build(rec, COMP[s[::-1]] if flag & 0x10 else s)
        out.append(rec)
    return np.concatenate(out)


@pytest.fixture(scope="module")
def case():
    ds = common.dataset("clip120k")
    here, elsewhere = split_secondary(ds["bam"])
    sm = np2.SecondarySeqs()
    for b in (here, elsewhere):
        sm.scan_ids(b)
    for b in (here, elsewhere):
        sm.scan_seqs(b)
    return ds, here, elsewhere, sm


def test_fill_matches_restatement_and_round_trips(case):
    ds, here, elsewhere, sm = case
    n_sec = sum(1 for r in records(here) if fields(r)[2] & 0x100)
    assert n_sec > 50 and sm.counts == (n_sec, n_sec)
    filled = sm.fill(here)
    assert np.array_equal(filled, expected_fill(here, elsewhere))
    # the recovered SEQ is the SEQ the original record had: only the secondary flag differs
    for a, b in zip(records(ds["bam"]), records(filled)):
        fa, fb = fields(a)[2], fields(b)[2]
        assert fb == fa or fb == fa | 0x100
        assert np.array_equal(unpack(a), unpack(b)) and len(a) == len(b)
    assert np.array_equal(sm.fill(elsewhere), elsewhere)  # no secondary record: untouched


def test_oracle_with_secondary_equals_original(case):
    ds, here, elsewhere, sm = case
    filled = sm.fill(here)
    tabs = common.oracle_tables(ds)
    want = O.Job(ds["contig"], ds["bam"], tabs, O.Opts(min_ctg_len=0), dump_iter=-1).consensus()
    got = O.Job(ds["contig"], filled, tabs, O.Opts(min_ctg_len=0, use_secondary=1), dump_iter=-1).consensus()
    common.assert_same("base", want[1], got[1])
    common.assert_same("pos", want[0], got[0])
    # without -S the secondary records are filtered: fewer reads, (generally) a different answer
    fewer = O.Job(ds["contig"], filled, tabs, O.Opts(min_ctg_len=0), dump_iter=0)
    assert len(fewer.reads()["t_s"]) < len(O.Job(ds["contig"], ds["bam"], tabs, O.Opts(min_ctg_len=0), dump_iter=0).reads()["t_s"])


def test_secondary_errors(case):
    ds, here, elsewhere, sm = case
    dup = np2.SecondarySeqs().scan_ids(here)
    dup.scan_seqs(elsewhere)
    with pytest.raises(np2.Np2Error) as e:  # assert!(seqs.insert(..).is_none()) secondary.rs:131
        dup.scan_seqs(elsewhere)
    assert e.value.code == -4
    with pytest.raises(np2.Np2Error):
        np2.SecondarySeqs().scan_ids(here[:-3])  # truncated record chain
    # a secondary record whose primary is nowhere: SEQ stays empty and the polish refuses it (reference: panic)
    lone = np2.SecondarySeqs().scan_ids(here).fill(here)
    assert np.array_equal(lone, here)
    with pytest.raises(O.OracleError):
        O.Job(ds["contig"], lone, common.oracle_tables(ds), O.Opts(min_ctg_len=0, use_secondary=1))


@pytest.mark.gpu
def test_gpu_with_secondary(ctx, case):
    ds, here, elsewhere, sm = case
    filled = sm.fill(here)
    oo, go = common.same_opts(use_secondary=1)
    opos, obase = O.Job(ds["contig"], filled, common.oracle_tables(ds), oo, dump_iter=-1).consensus()
    gpos, gbase = np2.polish_contig(ctx, ds["contig"], filled, common.gpu_tables(ctx, ds), go)
    common.assert_same("base", obase, gbase)
    common.assert_same("pos", opos, gpos)
    with pytest.raises(np2.Np2Error):
        np2.polish_contig(ctx, ds["contig"], here, common.gpu_tables(ctx, ds), go)  # SEQ-less secondary records


@pytest.mark.gpu
def test_cli_use_secondary(tmp_path, case):
    """nextPolish2 -S on a two-reference BAM: the primaries of the secondary reads live on the other reference."""
    ds, here, elsewhere, sm = case
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cli = os.path.join(root, "nextpolish2_b200", "nextPolish2")
    other = synth.genome(77, 30_000)
    order = np.argsort([int(np.frombuffer(r[8:12], "<i4")[0]) for r in records(elsewhere)], kind="stable")
    recs = list(records(elsewhere))
    elsewhere_sorted = np.concatenate([recs[i] for i in order])
    bam = str(tmp_path / "x.bam")
    synth.write_bam(bam, ["ctg", "other"], [len(ds["contig"]), len(other)], [here, elsewhere_sorted])
    fa = str(tmp_path / "x.fa")
    synth.write_fasta(fa, ["ctg", "other"], [ds["contig"], other], width=80)
    yaks = []
    for k in (21, 31):
        p = str(tmp_path / ("k%d.yak" % k))
        synth.write_yak(p, k, *ds["tables"][k])
        yaks.append(p)
    r = subprocess.run([cli, "-S", "-L", "100000", bam, fa] + yaks, capture_output=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    pos, base = O.Job(ds["contig"], ds["bam"], common.oracle_tables(ds), O.Opts(min_ctg_len=0), dump_iter=-1).consensus()
    want = O.format_fasta("ctg", pos, base) + O.format_fasta("other", np.arange(len(other), dtype=np.uint32), other)
    assert r.stdout == want
