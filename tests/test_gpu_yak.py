"""-m gpu: the HBM-resident yak table (K5) vs the oracle and vs the compiled reference's own file format."""
import os

import numpy as np
import pytest

import common
import oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("k", [21, 31, 51])
def test_lookup_matches_oracle(ctx, k):
    import nextpolish2_b200 as np2
    ds = common.dataset("hap300k")
    h, c = ds["tables"][k]
    ot = O.Table.from_arrays(k, h, c)
    gt = np2.Table.from_arrays(ctx, k, h, c)
    rng = np.random.default_rng(k)
    absent = rng.integers(0, 2**63, 200_000, dtype=np.uint64)
    near = h[rng.integers(0, len(h), 50_000)] ^ np.uint64(1 << 10)  # same sub-table, neighbouring tag
    flip = h[rng.integers(0, len(h), 50_000)] ^ np.uint64(1)        # same tag, other sub-table
    q = np.concatenate([h, absent, near, flip, np.array([0, 1023, 2**64 - 1], np.uint64)])
    for mc in (0, 1, 5, 40, 1023):
        common.assert_same("lookup k%d min%d" % (k, mc), ot.lookup(q, mc), gt.lookup(q, mc))
    got = gt.lookup(h, 0)
    common.assert_same("present keys return their count", c, got)
    assert len(gt) == len(h) and gt.k == k


@pytest.mark.parametrize("k", [21, 31, 51])
def test_golden_yak_file(ctx, k):
    """Tables written by the compiled reference `yak count` (tests/golden/*.yak): load through np2_yak_load and
    answer the committed (k-mer -> count) known answers produced by the reference's own yak_ch_get."""
    import json
    import nextpolish2_b200 as np2
    kat = json.load(open(os.path.join(GOLD, "yak_kat.json")))
    gt = np2.Table.load(ctx, os.path.join(GOLD, "k%d.yak" % k))
    e = kat["tables"][str(k)]
    assert len(gt) == e["n_keys"]
    seqs = [s.encode() for s in e["kmers"]]
    hashes = np.array([int(O.seq_hashes(s, k)[0]) for s in seqs], np.uint64)
    got = gt.lookup(hashes, 0)
    want = np.array([max(c, 0) for c in e["counts"]], np.uint16)  # yak_ch_get returns -1 for absent
    common.assert_same("golden counts k%d" % k, want, got)
    ks = gt.seq_kscore(seqs, 0)
    common.assert_same("golden kscore k%d" % k, want, ks)


@pytest.mark.parametrize("k", [21, 31, 51])
def test_seq_kscore(ctx, k):
    """iter2kmer + to_hash + get + min over arbitrary strings (N breaks windows, short strings score 0)."""
    import nextpolish2_b200 as np2
    ds = common.dataset("hap300k")
    h, c = ds["tables"][k]
    ot = O.Table.from_arrays(k, h, c)
    gt = np2.Table.from_arrays(ctx, k, h, c)
    hap = bytes(ds["haps"][0])
    rng = np.random.default_rng(7)
    seqs = []
    for i in range(3000):
        s = int(rng.integers(0, len(hap) - 400))
        L = int(rng.integers(1, 300))
        b = bytearray(hap[s:s + L])
        r = rng.random()
        if r < 0.3 and L > 3:
            b[int(rng.integers(0, L))] = ord("N")
        elif r < 0.5 and L > 3:
            b[int(rng.integers(0, L))] = ord("ACGT"[int(rng.integers(0, 4))])
        elif r < 0.55:
            b = bytearray(b.lower())
        seqs.append(bytes(b))
    seqs += [b"", b"A", b"ACGT" * 5, b"N" * 80]
    for mc in (1, 5, 45):
        want = []
        for s in seqs:
            hs = O.seq_hashes(s, k)
            want.append(int(ot.lookup(hs, mc).min()) if len(hs) else 0)
        common.assert_same("kscore k%d min%d" % (k, mc), np.array(want, np.uint16), gt.seq_kscore(seqs, mc))


def test_large_table_roundtrip(ctx):
    """Size-independent property at a size the oracle does not touch: every inserted key returns its count,
    a disjoint key set returns 0."""
    import nextpolish2_b200 as np2
    rng = np.random.default_rng(1)
    n = 20_000_000
    h = np.unique(rng.integers(0, 2**62, n, dtype=np.uint64))
    c = ((h >> np.uint64(13)) % np.uint64(1023) + np.uint64(1)).astype(np.uint16)
    gt = np2.Table.from_arrays(ctx, 31, h, c)
    common.assert_same("roundtrip", c, gt.lookup(h, 0))
    other = h ^ np.uint64(1 << 62)
    assert int(gt.lookup(other, 0).sum()) == 0
    flt = gt.lookup(h, 512)
    assert np.array_equal(flt, np.where(c >= 512, c, 0))


def test_table_image_adopt_clone(ctx):
    """Replication seam: a table adopted from the device image of another one, and a peer clone, answer
    exactly like the original (what bench.py's NVLink broadcast and the CLI's -g N rely on)."""
    import torch
    import nextpolish2_b200 as np2
    from nextpolish2_b200.shard import table_meta, _DeviceBytes
    ds = common.dataset("hap300k")
    h, c = ds["tables"][31]
    src = np2.Table.from_arrays(ctx, 31, h, c)
    meta = table_meta(src)
    ptr, nbytes, nb = src.image()
    assert meta == {"k": 31, "n_keys": len(h), "buckets_per_subtable": nb, "bytes": nbytes} and nbytes == src.device_bytes
    view = torch.as_tensor(_DeviceBytes(ptr, nbytes), device="cuda")  # zero-copy view, as the broadcast uses it
    buf = view.clone()                                                  # stands in for the received buffer
    torch.cuda.synchronize()
    rep = np2.Table.adopt(ctx, 31, len(h), nb, buf.data_ptr(), nbytes)
    del buf
    cl = src.clone(ctx)
    rng = np.random.default_rng(3)
    q = np.concatenate([h[::7], rng.integers(0, 2**63, 100_000, dtype=np.uint64)])
    want = src.lookup(q, 5)
    common.assert_same("adopted image", want, rep.lookup(q, 5))
    common.assert_same("peer clone", want, cl.lookup(q, 5))
    assert len(rep) == len(h) and rep.k == 31 and cl.device_bytes == nbytes
    with pytest.raises(np2.api.Np2Error):
        np2.Table.adopt(ctx, 31, len(h), nb + 1, ptr, nbytes)
