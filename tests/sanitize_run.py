"""Small end-to-end runs for compute-sanitizer (memcheck / racecheck / synccheck): a haploid and a diploid contig through
np2_job_create/upload/run twice each (the first run sizes everything exactly, the second one runs with speculative
capacities), compared with the oracle.  Kept small: the sanitizer slows kernels down 10-100x.

    compute-sanitizer --tool memcheck  python tests/sanitize_run.py > profiles/r02_sanitizer_memcheck.log
    compute-sanitizer --tool racecheck python tests/sanitize_run.py > profiles/r02_sanitizer_racecheck.log
`python tests/sanitize_run.py bgzf` runs only the contigs that come in as BGZF members (inflate + record-boundary kernels).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402

import oracle as O  # noqa: E402
import nextpolish2_b200 as np2  # noqa: E402
from nextpolish2_b200 import synth  # noqa: E402


def one(ctx, seed, L, het, depth, tandem=0.0, pinned=False):
    A = synth.genome(seed, L, tandem_frac=tandem)
    c = synth.make_contig(seed + 1, A, depth=depth, asm_err=3e-4, het=het, mean_len=6000, sd_len=1000, min_len=2000, threads=2)
    haps = [c["hap1"]] + ([c["hap2"]] if het > 0 else [])
    tabs = {k: synth.make_table(seed + 2, k, haps) for k in (21, 31)}
    gt = [np2.Table.from_arrays(ctx, k, *tabs[k]) for k in (21, 31)]
    oj = O.Job(A, c["bam"], [O.Table.from_arrays(k, *tabs[k]) for k in (21, 31)], O.Opts(min_ctg_len=0), dump_iter=-1)
    opos, obase = oj.consensus()
    pb = None
    bam = c["bam"]
    if pinned:  # page-locked records: K0 (the TMA gather of the CIGAR + SEQ spans) instead of the host compaction
        from nextpolish2_b200.api import PinnedBuffer
        pb = PinnedBuffer(bam)
        bam = pb.array
    job = np2.Job(ctx, A, bam, gt, np2.Opts(min_ctg_len=0)).upload()
    assert job.ingest_path == (1 if pinned else 2)
    for rnd in range(2):
        job.run(-1)
        gpos, gbase = job.consensus()
        assert np.array_equal(gbase, obase) and np.array_equal(gpos, opos), "GPU consensus differs from the oracle"
        assert np.array_equal(np.sort(job.dropped()), np.sort(oj.dropped()))
    st = job.stats()
    job.destroy()
    if pb is not None:
        pb.free()
    print("ok L=%d het=%g%s: %d regions, %d reads dropped, speculative passes %d, repeated %d" % (
        L, het, " (page-locked records)" if pinned else "", st["regions"], len(oj.dropped()), st["speculative_passes"],
        st["repeated_passes"]), flush=True)


def from_members(ctx, seed, L, het, depth, level):
    """the same contig from its BGZF members: k_bgzf_inflate, the record-boundary kernels, the device-to-device gather"""
    import tempfile
    A = synth.genome(seed, L)
    c = synth.make_contig(seed + 1, A, depth=depth, asm_err=3e-4, het=het, mean_len=6000, sd_len=1000, min_len=2000, threads=2)
    haps = [c["hap1"]] + ([c["hap2"]] if het > 0 else [])
    tabs = {k: synth.make_table(seed + 2, k, haps) for k in (21, 31)}
    gt = [np2.Table.from_arrays(ctx, k, *tabs[k]) for k in (21, 31)]
    oj = O.Job(A, c["bam"], [O.Table.from_arrays(k, *tabs[k]) for k in (21, 31)], O.Opts(min_ctg_len=0), dump_iter=-1)
    opos, obase = oj.consensus()
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "c.bam")
        synth.write_bam(path, ["ctg"], [L], [c["bam"]], level=level)
        buf = np.fromfile(path, np.uint8)
    po, pl, iz = np2.bgzf_members(buf)
    rec = bytes(c["bam"])
    got, _ = np2.bgzf_inflate(ctx, buf, po, pl, iz)
    skip = bytes(got).find(rec[:4096])
    assert skip > 0 and bytes(got)[skip:skip + len(rec)] == rec, "device inflate differs from the records written"
    job = np2.Job.from_bgzf(ctx, A, buf, po, pl, iz, skip, len(rec), gt, np2.Opts(min_ctg_len=0))
    assert job.ingest_path == 3
    job.upload().run(-1)
    gpos, gbase = job.consensus()
    assert np.array_equal(gbase, obase) and np.array_equal(gpos, opos), "GPU consensus differs from the oracle"
    job.destroy()
    print("ok L=%d het=%g from %d BGZF members (level %d): records inflated and walked on the device" % (L, het, len(po), level),
          flush=True)


def main():
    ctx = np2.Context(0)
    from_members(ctx, 51, 40_000, 0.003, 25, 1)
    from_members(ctx, 61, 30_000, 0.0, 20, 0)
    if len(sys.argv) > 1 and sys.argv[1] == "bgzf":  # only the ingest kernels (the rest has its logs already)
        return
    one(ctx, 11, 20_000, 0.0, 20)
    one(ctx, 21, 40_000, 0.004, 30)
    one(ctx, 31, 30_000, 0.002, 25, tandem=0.08)
    one(ctx, 41, 40_000, 0.004, 30, pinned=True)
    h = np.unique(np.random.default_rng(1).integers(0, 2**62, 200_000, dtype=np.uint64))
    t = np2.Table.from_arrays(ctx, 31, h, (h % 1000 + 1).astype(np.uint16))
    assert (t.lookup(h, 0) == (h % 1000 + 1)).all()
    print("ok table lookup", flush=True)


if __name__ == "__main__":
    main()
