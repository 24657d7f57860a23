"""-m gpu: the CUDA path through the C ABI vs the CPU oracle on the same seeded inputs, stage by stage and
end to end.  Everything here is integer/byte work: the bar is bit-exact equality."""
import hashlib

import numpy as np
import pytest

import common
import oracle as O

pytestmark = pytest.mark.gpu


def _run_both(ctx, name, dump_iter, ks=common.KS, **optkw):
    import nextpolish2_b200 as np2
    ds = common.dataset(name)
    oo, go = common.same_opts(**optkw)
    oj = O.Job(ds["contig"], ds["bam"], common.oracle_tables(ds, ks), oo, dump_iter=dump_iter)
    gt = common.gpu_tables(ctx, ds, ks)
    gj = np2.Job(ctx, ds["contig"], ds["bam"], gt, go).upload().run(dump_iter)
    return ds, oj, gj


@pytest.mark.parametrize("name", ["tiny20k", "hap300k", "clip120k", "dip600k"])
def test_stage_reads(ctx, name):
    """K1: filter + fill_with_cigar + trim(8) + AlignSeq::new + clip filter (main.rs:1758-1817)."""
    ds, oj, gj = _run_both(ctx, name, 0)
    common.assert_same_dict("reads", oj.reads(), gj.reads(), ["rec_idx", "t_s", "t_e", "blank", "nib_off", "nib"])


@pytest.mark.parametrize("name,it", [("tiny20k", 0), ("hap300k", 0), ("clip120k", 0), ("dip600k", 0), ("dip600k", 1),
                                      ("tandem200k", 0), ("tandem200k", 1)])
def test_stage_msa_dp(ctx, name, it):
    """K2 + K3: per-position 3-mer lists in Msa order with counts, besti, consensus, qv/cov flags."""
    ds, oj, gj = _run_both(ctx, name, it)
    om, gm = oj.msa(), gj.msa()
    common.assert_same_dict("msa", om, gm, ["off", "bases", "delta", "count"])
    # besti is only defined by the reference where a predecessor exists; dead entries keep 0 in both
    common.assert_same("msa.besti", om["besti"], gm["besti"])
    common.assert_same_dict("dp", oj.dp_consensus(), gj.dp_consensus())


@pytest.mark.parametrize("name,it", [("tiny20k", 0), ("hap300k", 1), ("clip120k", 0), ("dip600k", 0), ("dip600k", 1),
                                      ("deep80k", 0), ("tandem200k", 1)])
def test_stage_regions_candidates(ctx, name, it):
    """LQ regions, K4 candidates (order, seq, first-k k-mer hash) and K5 kscore."""
    ds, oj, gj = _run_both(ctx, name, it)
    common.assert_same_dict("regions", oj.regions(), gj.regions(), ["start", "end"])
    common.assert_same_dict("cand", oj.candidates(), gj.candidates(), ["roff", "order", "seq_off", "seq", "kmer", "kscore"])
    common.assert_same("regions.lable", oj.regions()["lable"], gj.regions()["lable"])
    common.assert_same("dropped", oj.dropped(), gj.dropped())


CASES = [
    ("tiny20k", common.KS, {}),
    ("hap300k", common.KS, {}),
    ("clip120k", common.KS, {}),
    ("dip600k", common.KS, {}),
    ("tandem200k", common.KS, {}),
    ("deep80k", common.KS, {}),
    ("dip600k", (21,), {"iter_count": 1}),
    ("clip120k", (21, 31, 51), {"iter_count": 3}),
    ("dip600k", (31,), {"use_all_reads": 1}),
    ("dip600k", common.KS, {"model": 1}),
    ("dip600k", common.KS, {"use_supplementary": 1, "min_map_qual": -1, "max_clip_len": 1000}),
    ("hap300k", common.KS, {"min_kmer_count": 40}),
    ("deep80k", (21, 51), {"max_indel_len": 1}),
    ("tandem200k", (21, 31, 51), {}),  # configs[3] in small: tandem repeats, three tables (k51 = bit-plane hash)
]


@pytest.mark.parametrize("name,ks,optkw", CASES, ids=[f"{c[0]}-k{'_'.join(map(str, c[1]))}-{'_'.join(c[2]) or 'default'}" for c in CASES])
def test_polish_fasta_identical(ctx, name, ks, optkw):
    """End to end through np2_polish_contig: FASTA bytes identical to the oracle (= reference -t 1)."""
    import nextpolish2_b200 as np2
    ds = common.dataset(name)
    oo, go = common.same_opts(**optkw)
    oj = O.Job(ds["contig"], ds["bam"], common.oracle_tables(ds, ks), oo, dump_iter=-1)
    opos, obase = oj.consensus()
    gpos, gbase = np2.polish_contig(ctx, ds["contig"], ds["bam"], common.gpu_tables(ctx, ds, ks), go)
    common.assert_same("consensus.base", obase, gbase)
    common.assert_same("consensus.pos", opos, gpos)
    fo = O.format_fasta("ctg1", opos, obase)
    fg = np2.format_fasta("ctg1", gpos, gbase)
    assert hashlib.sha256(fo).hexdigest() == hashlib.sha256(fg).hexdigest()
    # sanity property of the domain: the polish does something (consensus differs from the draft assembly)
    assert bytes(gbase) != bytes(ds["contig"])


@pytest.mark.parametrize("name,shift", [("clip120k", 0), ("dip600k", 5), ("exotic", 3)])
def test_pinned_records_device_gather(ctx, name, shift):
    """Page-locked record buffers take the K0 path (the device gathers SEQ over PCIe, QUAL never crosses the bus);
    pageable ones are compacted on the host.  Same reads, same consensus, at any buffer misalignment."""
    import exotic
    import nextpolish2_b200 as np2
    from nextpolish2_b200 import synth
    if name == "exotic":
        contig, bam = exotic.make()
        up = np.char.upper(contig.view("S1")).view(np.uint8)
        tabs = {k: synth.make_table(5, k, [up]) for k in (21, 31)}
    else:
        ds = common.dataset(name)
        contig, bam, tabs = ds["contig"], ds["bam"], ds["tables"]
    oo, go = common.same_opts()
    oj = O.Job(contig, bam, [O.Table.from_arrays(k, *tabs[k]) for k in (21, 31)], oo, dump_iter=0)
    gt = [np2.Table.from_arrays(ctx, k, *tabs[k]) for k in (21, 31)]
    pb = np2.PinnedBuffer(np.concatenate([np.zeros(shift, np.uint8), bam]))
    pc = np2.PinnedBuffer(contig)
    try:
        gj = np2.Job(ctx, pc.array, pb.array[shift:], gt, go)
        assert gj.ingest_path == 1
        gj.upload().run(0)
        common.assert_same_dict("reads", oj.reads(), gj.reads(), ["rec_idx", "t_s", "t_e", "blank", "nib_off", "nib"])
        opos, obase = oj.consensus()
        gpos, gbase = gj.consensus()
        common.assert_same("final.base", obase, gbase)
        common.assert_same("final.pos", opos, gpos)
        gj.destroy()
        pj = np2.Job(ctx, contig, bam, gt, go)
        assert pj.ingest_path == 2
        pj.destroy()
    finally:
        pb.free()
        pc.free()


def test_passthrough_and_errors(ctx):
    """Short contigs are echoed (main.rs:1727-1730); reference panics come back as error codes."""
    import nextpolish2_b200 as np2
    from nextpolish2_b200 import synth
    ds = common.dataset("tiny20k")
    gt = common.gpu_tables(ctx, ds)
    pos, base = np2.polish_contig(ctx, ds["contig"], ds["bam"], gt, np2.Opts())  # default -L 1000000
    assert bytes(base) == bytes(ds["contig"]) and pos[0] == 0 and pos[-1] == len(base) - 1
    seq = "ACGT" * 600
    bad = synth.bam_record(0, 10, [("M", 1200), ("N", 50), ("M", 1200)], seq)
    with pytest.raises(np2.Np2Error) as e:
        np2.polish_contig(ctx, ds["contig"], bad, gt, np2.Opts(min_ctg_len=0))
    assert e.value.code == -4 and "Unknown cigar" in str(e.value)
    ref = bytes(ds["contig"])
    r1 = synth.bam_record(0, 500, [("M", 2400)], ref[500:2900].decode())  # pushed reads: exact copies of the contig
    r0 = synth.bam_record(0, 100, [("M", 2400)], ref[100:2500].decode())
    with pytest.raises(np2.Np2Error) as e:
        np2.polish_contig(ctx, ds["contig"], np.concatenate([r1, r0]), gt, np2.Opts(min_ctg_len=0))
    assert "Unsorted" in str(e.value)
    with pytest.raises(O.OracleError):
        O.Job(ds["contig"], np.concatenate([r1, r0]), common.oracle_tables(ds), O.Opts(min_ctg_len=0))
    # a read that is never pushed (no 8-mer anchor) does not advance the sortedness cursor (main.rs:1814-1815)
    j1 = synth.bam_record(0, 500, [("M", 2400)], seq)
    mixed = np.concatenate([j1, r0])
    common.assert_same("unsorted-but-unpushed", O.Job(ds["contig"], mixed, common.oracle_tables(ds), O.Opts(min_ctg_len=0)).consensus()[1],
                       np2.polish_contig(ctx, ds["contig"], mixed, gt, np2.Opts(min_ctg_len=0))[1])


def test_no_reads(ctx):
    """A contig without any alignment: consensus == the contig's own codes (upper-cased ACGT)."""
    import nextpolish2_b200 as np2
    ds = common.dataset("tiny20k")
    empty = np.empty(0, np.uint8)
    contig = ds["contig"].copy()
    contig[100:200] |= 0x20  # lower-case stretch comes back upper-case (SEQ_NUM round trip)
    oj = O.Job(contig, empty, common.oracle_tables(ds), O.Opts(min_ctg_len=0))
    gpos, gbase = np2.polish_contig(ctx, contig, empty, common.gpu_tables(ctx, ds), np2.Opts(min_ctg_len=0))
    opos, obase = oj.consensus()
    common.assert_same("base", obase, gbase)
    common.assert_same("pos", opos, gpos)
    assert bytes(gbase) == bytes(ds["contig"])


@pytest.mark.parametrize("optkw", [{}, {"use_supplementary": 1, "min_map_qual": -1}, {"max_clip_len": 1000, "iter_count": 1},
                                   {"iter_count": 3, "use_all_reads": 1}])
def test_exotic_alignments(ctx, optkw):
    """IUPAC/N bases, lower-case and N stretches in the contig, H/S clip mixes, indels at read ends, long indels,
    =/X mixed with M, reads without an anchor, flagged records, zero-length ops (tests/exotic.py): every stage."""
    import exotic
    import nextpolish2_b200 as np2
    from nextpolish2_b200 import synth
    ref, blob = exotic.make()
    up = np.char.upper(ref.view("S1")).view(np.uint8)
    tabs = {k: synth.make_table(5, k, [up]) for k in (21, 31)}
    oo, go = common.same_opts(**optkw)
    for it in range(oo.iter_count):
        oj = O.Job(ref, blob, [O.Table.from_arrays(k, *tabs[k]) for k in (21, 31)], oo, dump_iter=it)
        gj = np2.Job(ctx, ref, blob, [np2.Table.from_arrays(ctx, k, *tabs[k]) for k in (21, 31)], go).upload().run(it)
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


@pytest.mark.parametrize("name", ["dip600k", "clip120k"])
def test_stage_pair_weights(ctx, name):
    """Stage seam VERDICT r01 asked for: the output of the pair loop of phase_reads_by_lqseqs (main.rs:953-992), one
    record per read pair with #agree + #differ * (2^32 - 1), exactly as it leaves K6 (k_edges_accum / k_edges_finish),
    against the oracle's per-region accumulation."""
    ds = common.dataset(name)
    oo, go = common.same_opts()
    oj = O.Job(ds["contig"], ds["bam"], common.oracle_tables(ds), oo, dump_iter=0)
    import nextpolish2_b200 as np2
    gj = np2.Job(ctx, ds["contig"], ds["bam"], common.gpu_tables(ctx, ds), go).upload().run(0)
    ok, ov = oj.pair_weights()
    gk, gv = gj.pair_weights()
    assert len(ok) > 1000
    common.assert_same("pair keys", ok, gk)
    common.assert_same("pair weights", ov, gv)
    assert (ok >> np.uint64(32) < (ok & np.uint64(0xFFFFFFFF))).all() and (np.diff(ok.astype(np.int64)) > 0).all()
    gj.destroy()


_LONG_DP_CHILD = r"""
import sys
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/oracle"); sys.path.insert(0, %(root)r + "/tests")
import numpy as np
import common, oracle as O, nextpolish2_b200 as np2
ctx = np2.Context(0)
for name in ("tandem200k", "dip600k", "deep80k"):
    ds = common.dataset(name)
    oo, go = common.same_opts()
    oj = O.Job(ds["contig"], ds["bam"], common.oracle_tables(ds), oo, dump_iter=0)
    gj = np2.Job(ctx, ds["contig"], ds["bam"], common.gpu_tables(ctx, ds), go).upload().run(0)
    common.assert_same_dict("msa", oj.msa(), gj.msa())          # besti of every Msa entry
    common.assert_same_dict("dp", oj.dp_consensus(), gj.dp_consensus())
    op, ob = oj.consensus(); gp, gb = gj.consensus()
    assert np.array_equal(ob, gb) and np.array_equal(op, gp)
    gj.destroy()
print("LONG-DP-OK")
"""


@pytest.mark.timeout(600)
def test_warp_cooperative_dp_matches_oracle():
    """K3 for long runs: with NP2_DP_LONG_WORK=0 EVERY run of multi-entry positions goes through the
    warp-cooperative kernel (entries of a position over the lanes, waves of equal b3.delta); besti of every Msa entry, the
    DP consensus and the final consensus must equal the oracle's on the tandem-repeat, diploid and deep sets."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, NP2_DP_LONG_WORK="0")
    r = subprocess.run([sys.executable, "-c", _LONG_DP_CHILD % {"root": root}], env=env, capture_output=True, text=True, timeout=550)
    assert r.returncode == 0 and "LONG-DP-OK" in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])


def test_speculative_capacities_across_unlike_contigs(ctx):
    """A context sizes the buffers of a pass from the LAST pass of the same kind, scaled by the contig length.  Contigs
    that are nothing like their predecessor (haploid -> diploid with ten times the events per base, long -> short,
    three iterations -> one) make those capacities wrong in both directions: every one of them has to be caught on the
    device (the pass is then repeated with exact sizes), never to end in a wrong consensus.  One context, resident
    hints, every result against the oracle."""
    import nextpolish2_b200 as np2
    seq = [("tiny20k", {}), ("dip600k", {"iter_count": 1}), ("tiny20k", {"iter_count": 1}), ("dip600k", {}),
           ("clip120k", {"iter_count": 2}), ("deep80k", {}), ("hap300k", {}), ("dip600k", {"iter_count": 2}), ("tiny20k", {})]
    own = np2.Context(0)  # fresh hints: the order above is the whole history
    repeated = 0
    oracle_cache = {}
    for name, optkw in seq:
        ds = common.dataset(name)
        oo, go = common.same_opts(**optkw)
        key = (name, tuple(sorted(optkw.items())))
        if key not in oracle_cache:
            oj = O.Job(ds["contig"], ds["bam"], common.oracle_tables(ds), oo, dump_iter=-1)
            oracle_cache[key] = oj.consensus() + (np.sort(oj.dropped()),)
        opos, obase, odrop = oracle_cache[key]
        tabs = common.gpu_tables(own, ds)
        job = np2.Job(own, ds["contig"], ds["bam"], tabs, go).upload().run(-1)
        gpos, gbase = job.consensus()
        common.assert_same("%s %r consensus.base" % (name, optkw), obase, gbase)
        common.assert_same("%s %r consensus.pos" % (name, optkw), opos, gpos)
        common.assert_same("%s %r dropped" % (name, optkw), odrop, np.sort(job.dropped()))
        repeated += job.stats()["repeated_passes"]
        job.destroy()
        for t in tabs:
            t.free()
    own.close()
    assert repeated > 0, "no pass was ever repeated: the sequence does not exercise a capacity overflow"
