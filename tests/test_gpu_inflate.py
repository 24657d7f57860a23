"""np2_bgzf_inflate (device: one warp per BGZF member) against zlib, through the C ABI."""
import random
import zlib

import numpy as np
import pytest

import bgzf as OB  # oracle/bgzf.py: member walk + zlib
import oracle as O
import nextpolish2_b200 as np2
from nextpolish2_b200 import synth

pytestmark = pytest.mark.gpu


def table(buf):
    ms = OB.members(buf)
    return (np.array([m[0] for m in ms], np.uint64), np.array([m[1] for m in ms], np.uint32),
            np.array([m[2] for m in ms], np.uint32))


@pytest.fixture(scope="module")
def bam_files(tmp_path_factory):
    d = tmp_path_factory.mktemp("bgzf")
    A = synth.genome(21, 400_000)
    c = synth.make_contig(22, A, depth=25, asm_err=1e-4, het=0.001, threads=4)
    out = {}
    for level in (0, 1, 6):
        path = str(d / ("t%d.bam" % level))
        synth.write_bam(path, ["ctg"], [len(A)], [c["bam"]], level=level)
        out[level] = np.fromfile(path, np.uint8)
    return out, bytes(c["bam"])


@pytest.mark.parametrize("level", [0, 1, 6])
def test_whole_file_matches_zlib(ctx, bam_files, level):
    files, records = bam_files
    buf = files[level]
    po, pl, iz = np2.bgzf_members(buf)
    opo, opl, oiz = table(buf)
    assert np.array_equal(po, opo) and np.array_equal(pl, opl) and np.array_equal(iz, oiz)
    want = OB.inflate_all(buf)
    got, ms = np2.bgzf_inflate(ctx, buf, po, pl, iz)
    assert bytes(got) == want
    assert ms > 0
    at = want.find(records[:4096])
    assert at > 0 and want[at:at + len(records)] == records  # the records sit behind the BAM header, as written


def test_subrange_pinned_buffers_and_member_subset(ctx, bam_files):
    files, _ = bam_files
    buf = files[1]
    po, pl, iz = np2.bgzf_members(buf)
    want = OB.inflate_all(buf)
    pin_in = np2.PinnedBuffer(buf)
    # the members 3 .. n-2 only, and a byte range that starts and ends inside a member (what a contig's records do)
    lo, hi = 3, len(po) - 2
    base = int(iz[:lo].astype(np.uint64).sum())
    total = int(iz[lo:hi].astype(np.uint64).sum())
    skip, n = 12345, total - 12345 - 6789
    pin_out = np2.PinnedBuffer(np.zeros(n, np.uint8))
    got, _ = np2.bgzf_inflate(ctx, pin_in, po[lo:hi], pl[lo:hi], iz[lo:hi], skip=skip, out_len=n, out=pin_out)
    assert bytes(got) == want[base + skip:base + skip + n]
    got2, _ = np2.bgzf_inflate(ctx, buf, po[lo:hi], pl[lo:hi], iz[lo:hi], skip=skip, out_len=n)  # pageable in and out
    assert bytes(got2) == bytes(got)
    pin_in.free()
    pin_out.free()


def test_every_block_type(ctx):
    """Members made with every zlib strategy / level: stored, fixed and dynamic blocks, several blocks per member,
    overlapping matches, empty members (the BGZF EOF marker)."""
    rng = random.Random(7)
    parts, datas = [], []

    def some(n):
        k = rng.randrange(5)
        if k == 0:
            return bytes(rng.getrandbits(8) for _ in range(n))
        if k == 1:
            return b"\xff" * n
        if k == 2:
            return (b"abc" * n)[:n]
        if k == 3:
            return bytes(rng.choice(b"\x11\x12\x14\x18\x21\x22\x24\x28\x41\x42\x44\x48\x81\x82\x84\x88") for _ in range(n))
        w = [bytes(rng.getrandbits(8) for _ in range(rng.randint(1, 30))) for _ in range(12)]
        b = b""
        while len(b) < n:
            b += rng.choice(w)
        return b[:n]
    for i in range(300):
        n = rng.choice([0, 1, 2, 17, 300, 4000, 30000, 65280, 65536])
        data = some(n)
        kw = dict(level=rng.choice([0, 1, 6, 9]), mem_level=rng.choice([1, 8, 9]),
                  strategy=rng.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE]))
        try:
            m = OB.make_member(data, **kw)
        except ValueError:  # incompressible: does not fit BSIZE
            data = data[:60000]
            m = OB.make_member(data, **kw)
        parts.append(m)
        datas.append(data)
    buf = np.frombuffer(b"".join(parts), np.uint8)
    po, pl, iz = np2.bgzf_members(buf)
    assert len(po) == 300
    got, _ = np2.bgzf_inflate(ctx, buf, po, pl, iz)
    assert bytes(got) == b"".join(datas)


def test_bad_member_is_an_error(ctx, bam_files):
    files, _ = bam_files
    buf = files[6].copy()
    po, pl, iz = np2.bgzf_members(buf)
    i = len(po) // 2
    wrong = iz.copy()
    wrong[i] -= 1  # ISIZE does not match the stream
    with pytest.raises(np2.Np2Error) as e:
        np2.bgzf_inflate(ctx, buf, po, pl, wrong)
    assert e.value.code == -4 and "BAM/SAM parsing failed!" in str(e.value) and ("member %d " % i) in str(e.value)
    buf[int(po[i]):int(po[i]) + 40] ^= 0x5A  # garbage block header / code lengths
    with pytest.raises(np2.Np2Error):
        np2.bgzf_inflate(ctx, buf, po, pl, iz)
    # the context is still usable
    got, _ = np2.bgzf_inflate(ctx, files[6], po, pl, iz)
    assert bytes(got) == OB.inflate_all(files[6])


@pytest.mark.parametrize("name", ["hap300k", "dip600k", "clip120k"])
def test_job_from_bgzf_members_equals_job_from_records(ctx, tmp_path, name):
    """np2_job_create_bgzf: members inflated on the device, record boundaries found there, only the record heads on
    the host — same consensus, same dropped reads, same ingest as the job made from the records themselves."""
    import common
    ds = common.dataset(name)
    A, rec = ds["contig"], np.ascontiguousarray(ds["bam"], np.uint8)
    path = str(tmp_path / "c.bam")
    synth.write_bam(path, ["ctg"], [len(A)], [rec], level=1)
    buf = np.fromfile(path, np.uint8)
    po, pl, iz = np2.bgzf_members(buf)
    skip = OB.inflate_all(buf).find(bytes(rec[:4096]))
    assert skip > 0
    tabs = common.gpu_tables(ctx, ds)
    opts = np2.Opts(min_ctg_len=0)
    a = np2.Job(ctx, A, rec, tabs, opts).upload().run()
    b = np2.Job.from_bgzf(ctx, A, buf, po, pl, iz, skip, len(rec), tabs, opts)
    assert b.ingest_path == 3
    b.upload().run()
    pa, ba = a.consensus()
    pb, bb = b.consensus()
    assert np.array_equal(ba, bb) and np.array_equal(pa, pb)
    assert np.array_equal(a.dropped(), b.dropped())
    oj = O.Job(A, rec, common.oracle_tables(ds), O.Opts(min_ctg_len=0))
    opos, obase = oj.consensus()
    assert np.array_equal(bb, obase) and np.array_equal(pb, opos)
    a.destroy()
    b.destroy()
    for t in tabs:
        t.free()


def test_job_from_bgzf_reports_bad_records(ctx, tmp_path):
    """a region that does not start at a record boundary: 'BAM/SAM parsing failed!' like the host path"""
    import common
    ds = common.dataset("tiny20k")
    A, rec = ds["contig"], np.ascontiguousarray(ds["bam"], np.uint8)
    path = str(tmp_path / "c.bam")
    synth.write_bam(path, ["ctg"], [len(A)], [rec], level=6)
    buf = np.fromfile(path, np.uint8)
    po, pl, iz = np2.bgzf_members(buf)
    skip = OB.inflate_all(buf).find(bytes(rec[:4096]))
    tabs = common.gpu_tables(ctx, ds)
    with pytest.raises(np2.Np2Error) as e:
        np2.Job.from_bgzf(ctx, A, buf, po, pl, iz, skip + 7, len(rec) - 7, tabs, np2.Opts(min_ctg_len=0))
    assert e.value.code == -4
    ok = np2.Job.from_bgzf(ctx, A, buf, po, pl, iz, skip, len(rec), tabs, np2.Opts(min_ctg_len=0)).upload().run()
    assert bytes(ok.consensus()[1]) == bytes(ds["haps"][0])
    ok.destroy()
    for t in tabs:
        t.free()


def test_job_from_bgzf_fallback_walks_the_records_on_the_host(tmp_path):
    """When the chunk join misses, the region is downloaded once and parsed by the host parser; the spans are still
    gathered from the device's records.  NP2_BGZF_HOST_PARSE=1 forces that path (own process: the knob is read once)."""
    import os
    import subprocess
    import sys
    code = """
import sys, numpy as np
sys.path[:0] = [%r, %r, %r]
import common, bgzf as OB, oracle as O, nextpolish2_b200 as np2
from nextpolish2_b200 import synth
ds = common.dataset("clip120k")
A, rec = ds["contig"], np.ascontiguousarray(ds["bam"], np.uint8)
synth.write_bam(%r, ["ctg"], [len(A)], [rec], level=1)
buf = np.fromfile(%r, np.uint8)
po, pl, iz = np2.bgzf_members(buf)
skip = OB.inflate_all(buf).find(bytes(rec[:4096]))
ctx = np2.Context(0)
tabs = common.gpu_tables(ctx, ds)
j = np2.Job.from_bgzf(ctx, A, buf, po, pl, iz, skip, len(rec), tabs, np2.Opts(min_ctg_len=0))
assert j.ingest_path == 3
j.upload().run()
pos, base = j.consensus()
opos, obase = O.Job(A, rec, common.oracle_tables(ds), O.Opts(min_ctg_len=0)).consensus()
assert np.array_equal(base, obase) and np.array_equal(pos, opos)
print("fallback ok")
"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = str(tmp_path / "f.bam")
    r = subprocess.run([sys.executable, "-c", code % (root, os.path.join(root, "oracle"), os.path.join(root, "tests"), path, path)],
                       capture_output=True, text=True, timeout=600, env=dict(os.environ, NP2_BGZF_HOST_PARSE="1"))
    assert r.returncode == 0 and "fallback ok" in r.stdout, r.stderr[-2000:]
