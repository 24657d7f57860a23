"""Host-side record parse (main.rs:1758-1771 filter + fill_with_cigar bookkeeping 386-440): the block_size chain is
walked speculatively from several byte offsets at once; whatever the guesses are, the result must be the one a
sequential reader produces.  Host only (np2_debug_parse needs no device)."""
import numpy as np
import pytest

import common
from nextpolish2_b200 import api, synth


def _parse(bam, tlen, threads, **kw):
    return api.debug_parse(bam, tlen, api.Opts(min_ctg_len=0, **kw), threads)


@pytest.mark.parametrize("name", ["tiny20k", "clip120k", "dip600k"])
def test_ranges_do_not_change_the_parse(name):
    ds = common.dataset(name)
    L = len(ds["contig"])
    serial = _parse(ds["bam"], L, 1)
    assert serial["fallback"] == 0 and serial["reads"] > 0 and serial["ops"] >= serial["reads"]
    for t in (2, 3, 7, 16, 64):
        got = _parse(ds["bam"], L, t)
        assert {k: got[k] for k in ("records", "reads", "ops", "columns", "digest")} == \
               {k: serial[k] for k in ("records", "reads", "ops", "columns", "digest")}, (t, got, serial)
    # the filter options reach the parse
    assert _parse(ds["bam"], L, 4, min_read_len=10**6)["reads"] == 0


def _decoy_record(ref, pos, n):
    """A real record whose QUAL bytes spell a chain of five well-formed little records (a wrong but plausible
    place to start walking)."""
    rec = synth.bam_record(0, pos, [("M", n)], ref[pos:pos + n].tobytes().decode()).copy()
    decoy = np.concatenate([synth.bam_record(0, 7, [("M", 40)], "ACGT" * 10, name="decoy") for _ in range(5)])
    assert len(decoy) < n - 64
    rec[len(rec) - n + 16:len(rec) - n + 16 + len(decoy)] = decoy  # inside QUAL (the last n bytes of the record)
    return rec


def test_decoy_chains_inside_qual_are_rejected():
    ds = common.dataset("tiny20k")
    ref = ds["contig"]
    L = len(ref)
    recs = [_decoy_record(ref, p, 2400) for p in range(0, L - 2400, 150)]
    bam = np.concatenate(recs)
    serial = _parse(bam, L, 1)
    assert serial["records"] == len(recs) and serial["reads"] == len(recs)
    missed = 0
    for t in (2, 5, 8, 16, 33):
        got = _parse(bam, L, t)
        missed += got["fallback"]
        assert got["digest"] == serial["digest"] and got["records"] == serial["records"], (t, got)
    assert missed > 0, "no byte range ever started on a decoy: the test does not exercise the fallback"


def test_first_error_in_file_order_wins():
    ds = common.dataset("tiny20k")
    ref = ds["contig"]
    L = len(ref)
    good = [synth.bam_record(0, p, [("M", 2400)], ref[p:p + 2400].tobytes().decode()) for p in range(0, L - 2400, 400)]
    bad_cigar = synth.bam_record(0, 300, [("M", 1200), ("N", 5), ("M", 1200)], "ACGT" * 600)
    for t in (1, 4, 9):
        with pytest.raises(api.Np2Error) as e:
            _parse(np.concatenate(good[:10] + [bad_cigar] + good[10:]), L, t)
        assert "Unknown cigar" in str(e.value)
        trunc = np.concatenate(good + [good[0][:100]])
        with pytest.raises(api.Np2Error) as e:
            _parse(trunc, L, t)
        assert "BAM/SAM parsing failed" in str(e.value)
        # both: the unknown CIGAR comes first in the file, a sequential reader never reaches the truncation
        with pytest.raises(api.Np2Error) as e:
            _parse(np.concatenate(good[:3] + [bad_cigar] + good[3:] + [good[0][:100]]), L, t)
        assert "Unknown cigar" in str(e.value)
    # a rejected record (MAPQ 0) may hold anything in its CIGAR (the reference filters before fill_with_cigar)
    lowq = synth.bam_record(0, 300, [("M", 1200), ("N", 5), ("M", 1200)], "ACGT" * 600, mapq=0)
    assert _parse(np.concatenate(good[:5] + [lowq] + good[5:]), L, 3)["reads"] == len(good)
    assert _parse(np.empty(0, np.uint8), L, 0) == {"records": 0, "reads": 0, "ops": 0, "columns": 0, "fallback": 0,
                                                   "digest": _parse(np.empty(0, np.uint8), L, 1)["digest"]}


def _expected_counts(bam, min_read_len=1000, min_map_len=500, min_map_fra=0.5, min_map_qual=1):
    """records / candidate reads / column-consuming ops / alignment columns from the pure-Python record reader and the
    record-level filter of main.rs:1758-1771 (tests/py_restatement.py); zero-length ops make no column and no op"""
    import py_restatement as P
    recs = P.records(bam)
    reads = ops = cols = 0
    for _tid, _pos, mapq, flag, cig, _seq in recs:
        rlen = sum(l for op, l in cig if op in (0, 1, 4, 7, 8, 5))
        span = sum(l for op, l in cig if op in (0, 2, 3, 7, 8))
        if flag & 4 or not cig or span == 0:
            span = 1
        frac = int(float(np.float32(rlen) * np.float32(min_map_fra)))
        if (flag & 0x404 or mapq <= min_map_qual or rlen <= min_read_len or flag & 0x900
                or span < max(min_map_len, frac)):
            continue
        reads += 1
        ops += sum(1 for op, l in cig if op in (0, 1, 2, 7, 8) and l > 0)
        cols += sum(l for op, l in cig if op in (0, 1, 2, 7, 8))
    return {"records": len(recs), "reads": reads, "ops": ops, "columns": cols}


def test_counts_match_an_independent_record_reader():
    import exotic
    ref, blob = exotic.make()  # flags, MAPQ 0/1, clips, N bases, zero-length ops, long indels
    sets = [(blob, len(ref))] + [(common.dataset(n)["bam"], len(common.dataset(n)["contig"])) for n in ("tiny20k", "clip120k")]
    for bam, L in sets:
        want = _expected_counts(bam)
        for t in (1, 3):
            got = _parse(bam, L, t)
            assert {k: got[k] for k in want} == want
    want = _expected_counts(blob, min_map_qual=-1, min_read_len=0)
    got = _parse(blob, len(ref), 2, min_map_qual=-1, min_read_len=0)
    assert {k: got[k] for k in want} == want and want["reads"] > _expected_counts(blob)["reads"]


FAST, SCALARS = 0x10000, 0x20000  # np2_debug_parse flags: the job path's parse / host-built ops left out of the digest
HEADS = 0x40000                   # the parse of np2_job_create_bgzf: record heads + offsets only


def _same_scalars(bam, L, **kw):
    keys = ("records", "reads", "ops", "columns", "digest")
    for t in (1, 3):
        a, b = _parse(bam, L, t | FAST, **kw), _parse(bam, L, t | SCALARS, **kw)
        assert {k: a[k] for k in keys} == {k: b[k] for k in keys}, (t, a, b)
        c = _parse(bam, L, t | HEADS, **kw)  # same records seen through their heads only
        assert {k: a[k] for k in keys} == {k: c[k] for k in keys}, (t, a, c)
    return a


def test_summed_cigar_parse_equals_the_op_building_parse():
    """np2_job_create only sums every CIGAR (the op records are expanded on the device); filter decisions, sizes and
    offsets must be the ones of the loop that builds the records, on real-looking data and on odd CIGARs."""
    import exotic
    ref, blob = exotic.make()
    _same_scalars(blob, len(ref))
    _same_scalars(blob, len(ref), min_map_qual=-1, min_read_len=0)
    for n in ("tiny20k", "clip120k", "dip600k"):
        ds = common.dataset(n)
        assert _same_scalars(ds["bam"], len(ds["contig"]))["reads"] > 0
    # random CIGARs over every op the reference knows, soft / hard clips anywhere, zero lengths
    rng = np.random.default_rng(5)
    L = 50_000
    recs = []
    for i in range(400):
        ops, q = [], 0
        # short CIGARs, and long ones mostly made of M I D = X (the parse sums those eight at a time)
        n_ops = int(rng.integers(1, 12)) if i % 2 else int(rng.integers(16, 90))
        for _ in range(n_ops):
            o = "MIDSH=X"[int(rng.integers(0, 7))] if i % 2 or rng.random() < 0.03 else "MID=X"[int(rng.integers(0, 5))]
            l = int(rng.integers(0, 400)) if rng.random() < 0.98 else int(rng.integers(1 << 24, 1 << 25))
            ops.append((o, l))
            if o in "MIS=X":
                q += l
        extra = int(rng.integers(0, 3))  # SEQ may be longer than the CIGAR consumes
        if q > 1 << 22:  # keep the test's memory small: a huge op makes the record fail the l_seq check in both parses
            q = 1000
        recs.append(synth.bam_record(0, int(rng.integers(0, 40_000)), ops, "ACGT"[i % 4] * (q + extra), mapq=int(rng.integers(0, 61))))
    bam = np.concatenate(recs)
    # records that would make the reference panic are taken out one by one (both parses must name the same one)
    while True:
        try:
            got = _same_scalars(bam, L, min_read_len=0, min_map_len=0, min_map_qual=-1)
            break
        except api.Np2Error as e:
            msgs = []
            for flags in (FAST, SCALARS):
                with pytest.raises(api.Np2Error) as ee:
                    _parse(bam, L, 1 | flags, min_read_len=0, min_map_len=0, min_map_qual=-1)
                msgs.append(str(ee.value))
            assert msgs[0] == msgs[1] == str(e)
            # drop the first record either parse rejects: find it by bisection over prefixes
            lo, hi = 0, len(recs)
            while hi - lo > 1:
                mid = (lo + hi) // 2
                try:
                    _parse(np.concatenate(recs[:mid]), L, 1 | FAST, min_read_len=0, min_map_len=0, min_map_qual=-1)
                    lo = mid
                except api.Np2Error:
                    hi = mid
            del recs[lo]
            bam = np.concatenate(recs)
    assert got["records"] >= 200 and got["reads"] > 100


def test_summed_cigar_parse_reports_the_same_errors():
    ds = common.dataset("tiny20k")
    ref = ds["contig"]
    L = len(ref)
    good = [synth.bam_record(0, p, [("M", 2400)], ref[p:p + 2400].tobytes().decode()) for p in range(0, L - 2400, 400)]
    cases = {
        "Unknown cigar": synth.bam_record(0, 300, [("M", 1200), ("N", 5), ("M", 1200)], "ACGT" * 600),
        "more query bases": synth.bam_record(0, 300, [("M", 1200), ("I", 10), ("M", 1200)], "ACGT" * 600),
        "past the end": synth.bam_record(0, L - 1000, [("M", 1200), ("D", 3), ("M", 1200)], "ACGT" * 600),
        "outside the contig": synth.bam_record(0, L + 5, [("M", 2400)], "ACGT" * 600),
    }
    for msg, bad in cases.items():
        for flags in (FAST, 0, HEADS):
            with pytest.raises(api.Np2Error) as e:
                _parse(np.concatenate(good[:4] + [bad] + good[4:]), L, 2 | flags)
            assert msg in str(e.value), (msg, flags, str(e.value))
    # a trailing soft clip may run past SEQ without complaint (only aligned ops are checked), in both parses
    tail = synth.bam_record(0, 300, [("M", 2400), ("S", 50)], "ACGT" * 600)
    _same_scalars(np.concatenate(good[:4] + [tail] + good[4:]), L)
