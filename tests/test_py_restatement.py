"""The C++ oracle against an independent pure-Python restatement of the same reference code (tests/py_restatement.py):
reads after trim / clip filter, the per-position 3-mer lists in Msa order with counts and back pointers, the DP
consensus with its qv / coverage flags, the LQ regions and the candidate alleles of every region (read order, string,
first-k k-mer hash, the 60 cap, the monotone region cursor) must all be identical; so must the reads blanked by
the phasing step, the Louvain result on synthetic read graphs and the final consensus (positions and bases) of the
whole path with one or two k-mer tables.  CPU only."""
import gzip
import os

import numpy as np
import pytest

import common
import exotic
import oracle as O
import py_restatement as P
from nextpolish2_b200 import synth


def compare(contig, bam, table=None, **optkw):
    tseq = bytes(contig).decode()
    if table is None:
        table = (np.array([1], np.uint64), np.array([9], np.uint16))
    # iter_count = 2 and the dump of iteration 0: the non-final iteration runs retrieve_kmer_count + mark_hete_lqseqs
    oj = O.Job(contig, bam, [O.Table.from_arrays(21, *table)], O.Opts(min_ctg_len=0, iter_count=2, **optkw), dump_iter=0)
    als, rec_idx = P.ingest(tseq, bam, **{k: v for k, v in optkw.items() if k in ("max_clip_len", "min_map_qual", "use_supplementary")})
    r = oj.reads()
    assert list(r["rec_idx"]) == rec_idx
    assert [a.aln_t_s for a in als] == list(r["t_s"]) and [a.aln_t_e for a in als] == list(r["t_e"])
    assert [0 if a.align_bases else 1 for a in als] == list(r["blank"])
    nib = b"".join(bytes(a.align_bases) for a in als[1:])  # the ref read (alignseq 0) is implicit in the dump
    got = bytes(r["nib"])
    assert got == nib or got == b"".join(bytes(a.align_bases) for a in als)
    msas = P.build_msas(len(tseq), als)
    best = P.dp(msas)
    m = oj.msa()
    assert list(m["off"]) == list(np.cumsum([0] + [len(x) for x in msas]))
    assert list(m["bases"]) == [k.bases for x in msas for k in x]
    assert list(m["delta"]) == [k.delta for x in msas for k in x]
    assert list(m["count"]) == [k.count for x in msas for k in x]
    assert list(m["besti"]) == [k.besti for x in msas for k in x]
    cns, regions = P.backtrack(msas, best)
    d = oj.dp_consensus()
    assert list(d["pos"]) == [c[0] for c in cns]
    assert bytes(d["base"]).decode() == "".join(c[1] for c in cns)
    assert list(d["flags"]) == [c[2] for c in cns]
    reg = oj.regions()
    assert list(zip(reg["start"], reg["end"])) == regions
    if regions:  # candidates in read order with the 60 cap: order, string, hash of the first-k k-mer (k = 21)
        cand = P.candidates(als, regions, 21)
        c = oj.candidates()
        assert list(c["roff"]) == list(np.cumsum([0] + [len(x) for x in cand]))
        flat = [x for r in cand for x in r]
        assert list(c["order"]) == [x[0] for x in flat]
        assert bytes(c["seq"]).decode() == "".join(x[1] for x in flat)
        assert list(c["seq_off"]) == list(np.cumsum([0] + [len(x[1]) for x in flat]))
        assert [int(v) for v in c["kmer"]] == [x[2] for x in flat]
        # k-mer scores (retrieve_kmer_count) and heterozygous regions (fill_order_stat + mark_hete_lqseqs, which also
        # zeroes the scores of the minor alleles)
        tab = {int(h) >> 10: int(n) for h, n in zip(*table)}
        ks = P.kscores(cand, tab, 21)
        assert [int(v) for v in c["kscore"]] == [x for r in ks for x in r]  # (the oracle dumps them before mark_hete)
        hete = P.mark_hete(cand, ks)
        assert [bool(l & 0x40) for l in reg["lable"]] == hete
        return len(als), len(regions), sum(hete)
    return len(als), len(regions), 0


def test_synthetic_haploid_and_diploid():
    ds = common.dataset("tiny20k")
    n, nreg, _ = compare(ds["contig"], ds["bam"], table=ds["tables"][21])
    assert n > 30 and nreg > 0
    A = synth.genome(41, 12_000)
    c = synth.make_contig(42, A, depth=25, asm_err=1e-3, het=0.004, mean_len=4000, sd_len=600, min_len=1500,
                          frac_clip=0.05, frac_lowq=0.03, frac_supp=0.03, eqx=True, read_err=0.006, threads=2)
    tab = synth.make_table(43, 21, [c["hap1"], c["hap2"]])
    n, nreg, nhete = compare(A, c["bam"], table=tab)
    assert nreg > 5 and nhete > 3  # real heterozygous sites are recognised by both
    compare(A, c["bam"], table=tab, use_supplementary=1, min_map_qual=-1, max_clip_len=1000)


def test_exotic_alignments():
    """IUPAC / N bases, lower case, clips, indels at the ends, long indels, reads without an anchor (tests/exotic.py)"""
    ref, blob = exotic.make()
    compare(ref, blob)
    compare(ref, blob, max_clip_len=1000)


def test_real_reads():
    """the bundled 40 kb contig with the real HiFi reads aligned to it (configs[0] data, ~74x): every stage, then the
    whole path with both tables"""
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c1_40k")
    contig = np.frombuffer(gzip.open(os.path.join(d, "contig.bin.gz")).read(), np.uint8)
    bam = np.frombuffer(gzip.open(os.path.join(d, "records.bin.gz")).read(), np.uint8)
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_c1 import read_yak
    _, h, cnt = read_yak(os.path.join(d, "k21.yak"))
    n, nreg, nhete = compare(contig, bam, table=(h, cnt))
    assert n > 100 and nreg > 500 and nhete > 100
    _, h31, cnt31 = read_yak(os.path.join(d, "k31.yak"))
    ndrop, _ = compare_full(contig, bam, {21: (h, cnt), 31: (h31, cnt31)})
    assert ndrop > 50  # real heterozygosity: reads of the other haplotype are blanked before the final iteration


def compare_full(contig, bam, tables, **optkw):
    """the whole path: reads blanked by the phasing of iteration 0 and the final consensus (positions + bases)"""
    tseq = bytes(contig).decode()
    ots = [O.Table.from_arrays(k, h, c) for k, (h, c) in tables.items()]
    oj = O.Job(contig, bam, ots, O.Opts(min_ctg_len=0, **optkw), dump_iter=0)
    want_pos, want_base = oj.consensus()
    ptabs = [(k, {int(h) >> 10: int(n) for h, n in zip(*t)}) for k, t in tables.items()]
    kw = {k: v for k, v in optkw.items() if k in ("iter_count", "max_indel_len", "min_kmer_count")}
    cns, dropped = P.polish(tseq, bam, ptabs, asref=optkw.get("model", 0) == 0,
                            use_all_reads=bool(optkw.get("use_all_reads", 0)), **kw)
    # the oracle's list accumulates over the non-final iterations (a blanked read cannot be dropped again)
    assert sorted(int(x) for x in oj.dropped()) == sorted(x for d in dropped for x in d)
    assert [p for p, _ in cns] == list(want_pos)
    assert "".join(b for _, b in cns) == bytes(want_base).decode()
    changed = "".join(b for _, b in cns) != tseq
    return len(dropped[0]) if dropped else 0, changed


def test_full_path_diploid():
    """seed choice, Louvain phasing, consensus patching and the k-mer re-check with two tables (k = 21, 31)"""
    A = synth.genome(51, 12_000)
    c = synth.make_contig(52, A, depth=25, asm_err=1e-3, het=0.004, mean_len=4000, sd_len=600, min_len=1500,
                          read_err=0.006, threads=2)
    tabs = {k: synth.make_table(53, k, [c["hap1"], c["hap2"]]) for k in (21, 31, 51)}
    compare_full(A, c["bam"], tabs)  # the third re-check uses the bit-plane k-mers of k >= 32 (kmer.rs:288-309)
    del tabs[51]
    ndrop, changed = compare_full(A, c["bam"], tabs)
    assert ndrop > 3 and changed  # reads of the other haplotype are blanked, assembly errors are corrected
    for kw in ({"model": 1}, {"use_all_reads": 1}, {"iter_count": 1}, {"iter_count": 3}, {"max_indel_len": 0}):
        compare_full(A, c["bam"], tabs, **kw)
    compare_full(A, c["bam"], {21: tabs[21]})


def test_full_path_haploid_and_exotic():
    ds = common.dataset("tiny20k")
    compare_full(ds["contig"], ds["bam"], {21: ds["tables"][21], 31: ds["tables"][31]})
    compare_full(ds["contig"], ds["bam"], {51: ds["tables"][51], 21: ds["tables"][21]})  # bit-plane k-mers, k >= 32
    ref, blob = exotic.make()
    compare_full(ref, blob, {21: (np.array([1], np.uint64), np.array([9], np.uint16))})


@pytest.mark.parametrize("model,use_all", [(0, False), (1, False), (0, True)])
def test_louvain_phasing_on_read_graphs(model, use_all):
    """louvain.rs through both restatements on synthetic two-haplotype read graphs (tests/test_phase_host.py), the
    community that falls apart again included; the Python side replays the per-site +1 / -1 updates one at a time"""
    import test_phase_host as T
    D = (1 << 32) - 1
    graphs = [T.make_graph(seed=1, n_reads=60), T.make_graph(seed=21, n_reads=250, span=20, noise=0.08),
              T.make_graph(seed=22, n_reads=200, span=15, noise=0.3, gap_every=40),
              T.make_graph(seed=23, n_reads=150, span=20, ref=False), T.gadget()]
    k2, v2 = T.make_graph(seed=24, n_reads=80, span=10)
    gk, gv = T.gadget(base=100)
    graphs.append((np.concatenate([k2, gk]), np.concatenate([v2, gv])))
    for keys, vals in graphs:
        pairs = []
        for k, v in zip(keys.tolist(), vals.tolist()):
            differ = (v + (1 << 31)) >> 32
            pairs.append((k >> 32, k & D, v - differ * D, differ))
            assert pairs[-1][2] >= 0
        want = O.debug_phase(keys, vals, model, use_all)
        assert P.phase_from_pairs(pairs, model == 0, use_all) == [int(x) for x in want]


def test_real_reads_whole_bundled_contig():
    """configs[0] in full: the whole bundled contig (100 kb), every read, the unmodified yak tables (git-ignored fixture
    built by tests/golden/make_c1.py --full at build() time)"""
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "_c1")
    if not os.path.exists(os.path.join(d, "records.bin")):
        pytest.skip("fixture _c1 not built (tests/golden/make_c1.py needs /root/reference)")
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_c1 import read_yak
    contig = np.fromfile(os.path.join(d, "contig.bin"), np.uint8)
    bam = np.fromfile(os.path.join(d, "records.bin"), np.uint8)
    tabs = {k: read_yak(os.path.join(d, "k%d.yak" % k))[1:] for k in (21, 31)}
    n, nreg, nhete = compare(contig, bam, table=tabs[21])
    assert n > 300 and nreg > 2000 and nhete > 300
    ndrop, changed = compare_full(contig, bam, tabs)
    assert ndrop > 150 and changed  # (-m, -r and -i 3 agree as well; left out of the suite for time)
