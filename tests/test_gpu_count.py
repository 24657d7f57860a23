"""`yak count` on the device (np2_count_*) against the compiled reference (oracle/_ref/yak, built from
/root/reference/yak) and against the committed tables that binary wrote (tests/golden/k{21,31,51}.yak).

What must agree is the SET of (hash, count) pairs of every sub-table; the order of keys inside a sub-table is the
iteration order of yak's hash table and carries no meaning (yak_ch_restore / KmerInfo::retrieve_kmers re-insert)."""
import os
import subprocess

import numpy as np
import pytest

import oracle as O
from nextpolish2_b200 import synth

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
YAK = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "yak")


def read_yak(path):
    """-> k, {sub-table: sorted array of file keys}"""
    raw = np.fromfile(path, np.uint8)
    assert bytes(raw[:4]) == b"YAK\2"
    k, pre, bits = [int(x) for x in np.frombuffer(raw[4:16], "<u4")]
    assert pre == 10 and bits == 10
    off, subs = 16, []
    for b in range(1 << pre):
        cap, size = [int(x) for x in np.frombuffer(raw[off:off + 8], "<u4")]
        assert size == 0 or size <= cap
        subs.append(np.sort(np.frombuffer(raw[off + 8:off + 8 + 8 * size], "<u8")))
        off += 8 + 8 * size
    assert off == len(raw)
    return k, subs


def same_tables(a, b):
    ka, sa = read_yak(a)
    kb, sb = read_yak(b)
    assert ka == kb
    for x in range(1024):
        assert np.array_equal(sa[x], sb[x]), "sub-table %d differs" % x


def golden_reads(tmp_path):
    """the reads tests/golden/make_golden.py counted (same seeds, same writer)"""
    G = synth.genome(20260001, 3000)
    fa = str(tmp_path / "sr.fa")
    synth.write_short_reads(fa, 7, [G], depth_each=30.0, rlen=150, sub_rate=0.002)
    return fa, [s for _, s in synth.read_fasta(fa)]


@pytest.mark.parametrize("k", [21, 31, 51])
def test_count_matches_committed_reference_tables(ctx, tmp_path, k):
    import nextpolish2_b200 as np2
    fa, reads = golden_reads(tmp_path)
    c = np2.Counter(ctx, k).add(reads)
    out = str(tmp_path / ("gpu_k%d.yak" % k))
    tab = c.finish(min_count=1, dump_path=out, table=True)
    same_tables(out, os.path.join(HERE, "golden", "k%d.yak" % k))
    distinct, total = c.distinct
    assert total == sum(max(0, len(r) - k + 1) for r in reads)
    # the staged table answers like the oracle's table loaded from the reference's file
    ot = O.Table.load(os.path.join(HERE, "golden", "k%d.yak" % k))
    hs = np.unique(np.concatenate([O.seq_hashes(r, k) for r in reads[:50]]))
    rng = np.random.default_rng(k)
    q = np.concatenate([hs, rng.integers(0, 2**62, 500, dtype=np.uint64)])
    for mc in (1, 5, 40):
        assert np.array_equal(tab.lookup(q, mc), ot.lookup(q, mc))
    assert len(ot) == distinct
    c.close()


@pytest.mark.skipif(not os.path.exists(YAK), reason="oracle/_ref/yak not built")
@pytest.mark.parametrize("k,bloom", [(21, 0), (31, 0), (21, 24), (31, 20), (51, 22)])
def test_count_matches_live_reference(ctx, tmp_path, k, bloom):
    """Reads with N, lower-case, reads shorter than k, deep coverage (counts above 1023 saturate), several add() calls;
    with -b the reference makes two passes over the same file and keeps counts >= 2."""
    import nextpolish2_b200 as np2
    rng = np.random.default_rng(100 + k + bloom)
    G = synth.genome(31, 4000)
    g = bytes(G)
    reads = []
    for i in range(6000):
        s = int(rng.integers(0, len(g) - 150))
        r = bytearray(g[s:s + int(rng.integers(10, 151))])
        if i % 7 == 0 and len(r) > 30:
            r[int(rng.integers(0, len(r)))] = ord("N")
        if i % 11 == 0:
            r = bytearray(bytes(r).lower())
        if i % 5 == 0:
            r[int(rng.integers(0, len(r)))] = ord("ACGT"[int(rng.integers(0, 4))])
        reads.append(bytes(r))
    reads += [g[100:400]] * 1500  # more than 1023 copies of the same k-mers
    fa = str(tmp_path / "r.fa")
    with open(fa, "wb") as f:
        for i, r in enumerate(reads):
            f.write(b">r%d\n%s\n" % (i, r))
    ref = str(tmp_path / "ref.yak")
    cmd = [YAK, "count", "-k", str(k), "-t", "2", "-o", ref] + (["-b", str(bloom), fa, fa] if bloom else [fa])
    subprocess.check_call(cmd, stderr=subprocess.DEVNULL)
    c = np2.Counter(ctx, k)
    for part in (reads[:1000], reads[1000:4000], reads[4000:]):
        c.add(part)
    out = str(tmp_path / "gpu.yak")
    c.finish(min_count=2 if bloom else 1, dump_path=out)
    same_tables(out, ref)
    _, subs = read_yak(out)
    assert max(int((s & np.uint64(1023)).max()) for s in subs if len(s)) == 1023
    c.close()
    # the reference's own reader accepts our file: round trip through the library's loader
    t = np2.Table.load(ctx, out)
    assert t.k == k


def test_count_then_polish_uses_the_same_table(ctx, tmp_path):
    """k-mer tables counted here drive the same polish as tables counted by writing + loading a yak file."""
    import nextpolish2_b200 as np2
    import common
    ds = common.dataset("tiny20k")
    sr = str(tmp_path / "sr.fa")
    synth.write_short_reads(sr, 5, ds["haps"], depth_each=40.0, rlen=150, sub_rate=0.001)
    reads = [s for _, s in synth.read_fasta(sr)]
    tabs, files = [], []
    for k in (21, 31):
        c = np2.Counter(ctx, k).add(reads)
        p = str(tmp_path / ("k%d.yak" % k))
        tabs.append(c.finish(min_count=2, dump_path=p, table=True))
        files.append(p)
        c.close()
    opts = np2.Opts(min_ctg_len=0)
    a = np2.polish_contig(ctx, ds["contig"], ds["bam"], tabs, opts)
    b = np2.polish_contig(ctx, ds["contig"], ds["bam"], [np2.Table.load(ctx, p) for p in files], opts)
    o = O.Job(ds["contig"], ds["bam"], [O.Table.load(p) for p in files], O.Opts(min_ctg_len=0), dump_iter=-1).consensus()
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[1], o[1]) and np.array_equal(a[0], o[0])
    assert bytes(a[1]) == bytes(ds["haps"][0])


@pytest.mark.skipif(not os.path.exists(YAK), reason="oracle/_ref/yak not built")
def test_cli_count_subcommand(tmp_path):
    """`nextPolish2 count` (FASTQ.gz in, .yak out) against `yak count` on the same file, with and without -b."""
    import gzip
    cli = os.path.join(os.path.dirname(HERE), "nextpolish2_b200", "nextPolish2")
    G = synth.genome(77, 5000)
    fa = str(tmp_path / "sr.fa")
    synth.write_short_reads(fa, 3, [G], depth_each=25.0, rlen=150, sub_rate=0.003)
    fq = str(tmp_path / "sr.fq.gz")
    with gzip.open(fq, "wb") as f:
        for name, s in synth.read_fasta(fa):
            f.write(b"@" + name.encode() + b" extra\n" + s + b"\n+\n" + b"I" * len(s) + b"\n")
    for k, bloom in ((21, 0), (31, 22)):
        ref, out = str(tmp_path / "ref.yak"), str(tmp_path / "out.yak")
        subprocess.check_call([YAK, "count", "-k", str(k), "-o", ref] + (["-b", str(bloom), fq, fq] if bloom else [fq]),
                              stderr=subprocess.DEVNULL)
        r = subprocess.run([cli, "count", "-k", str(k), "-o", out] + (["-b", str(bloom), fq, fq] if bloom else [fq]),
                           capture_output=True, timeout=300)
        assert r.returncode == 0, r.stderr
        same_tables(out, ref)
