"""CPU: pins the oracle's k-mer hash / yak-table restatement (kmer.rs) against known answers produced by the
COMPILED reference C code (tests/golden/yak_kat.json, k*.yak; generator: tests/golden/make_golden.py) and, when
oracle/_ref is present, against the reference library directly."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import common
import oracle as O
from nextpolish2_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
REF = os.path.join(os.path.dirname(HERE), "oracle", "_ref")
KAT = json.load(open(os.path.join(GOLD, "yak_kat.json")))


def test_hash_known_answers():
    L = O.lib()
    for key, mask, want in KAT["hash64"]:
        assert L.np2o_yak_hash64(key, mask) == want
    for key, want in KAT["hash64_64"]:
        assert L.np2o_yak_hash64_64(key) == want
    for x, want in KAT["hash_long"]:
        arr = (C.c_uint64 * 4)(*x)
        assert L.np2o_yak_hash_long(arr) == want


@pytest.mark.parametrize("k", [21, 31, 51])
def test_table_known_answers(k):
    """oracle: np2o_table_load(reference-written file) + iter2kmer/to_hash + lookup == reference yak_ch_get."""
    t = O.Table.load(os.path.join(GOLD, "k%d.yak" % k))
    e = KAT["tables"][str(k)]
    assert len(t) == e["n_keys"] and t.k == k
    hashes = np.array([int(O.seq_hashes(s.encode(), k)[0]) for s in e["kmers"]], np.uint64)
    want = np.array([max(c, 0) for c in e["counts"]], np.uint16)
    common.assert_same("counts", want, t.lookup(hashes, 0))
    # min_count semantics (kmer.rs:158-166): count >= min kept, else 0
    for mc in (5, 20):
        common.assert_same("min%d" % mc, np.where(want >= mc, want, 0), t.lookup(hashes, mc))
    t.set_stream_scan(True)  # the reference's full-file streaming form gives the same answers
    common.assert_same("stream", want, t.lookup(hashes, 0))


def test_canonical_and_window_rules():
    k = 21
    s = b"ACGTTGCATGCATGCATTAGCGATCGATCGATTAGC"
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    fwd = O.seq_hashes(s, k)
    rev = O.seq_hashes(s.translate(comp)[::-1], k)
    assert len(fwd) == len(s) - k + 1 and np.array_equal(fwd, rev[::-1])  # strand-canonical
    assert np.array_equal(O.seq_hashes(s.lower(), k), fwd)               # SEQ_NUM folds case
    n = bytearray(s)
    n[25] = ord("N")                                                       # a non-ACGT base resets the window
    assert len(O.seq_hashes(bytes(n), k)) == max(0, 25 - k + 1) + max(0, len(s) - 26 - k + 1)
    assert len(O.seq_hashes(b"ACGT", k)) == 0


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "libyakref.so")), reason="oracle/_ref not built")
@pytest.mark.parametrize("k", [21, 31, 51])
def test_against_compiled_reference_live(k, tmp_path):
    """Everything in a freshly counted table: reference yak_ch_get == oracle lookup, for every key of the synthetic
    genome and for random absent keys; also validates the synthetic table writer's file layout."""
    lib = C.CDLL(os.path.join(REF, "libyakref.so"))
    lib.yakref_restore.restype = C.c_void_p
    lib.yakref_restore.argtypes = [C.c_char_p]
    lib.yakref_get.argtypes = [C.c_void_p, C.c_uint64]
    G = synth.genome(99, 20000)
    h, c = synth.make_table(5, k, [G], mean_count=30.0, keep_min=1)
    path = str(tmp_path / "t.yak")
    synth.write_yak(path, k, h, c)
    ref = lib.yakref_restore(path.encode())
    assert ref
    t = O.Table.load(path)
    assert len(t) == len(h)
    oh = O.seq_hashes(G, k)
    assert set(oh.tolist()) == set(h.tolist())
    got = t.lookup(oh, 0)
    want = np.array([max(lib.yakref_get(ref, int(x)), 0) for x in oh[:4000]], np.uint16)
    common.assert_same("live", want, got[:4000])
    rnd = np.random.default_rng(3).integers(0, 2**63, 2000, dtype=np.uint64)
    want = np.array([max(lib.yakref_get(ref, int(x)), 0) for x in rnd], np.uint16)
    common.assert_same("absent", want, t.lookup(rnd, 0))


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "yak")), reason="oracle/_ref not built")
def test_synth_table_matches_real_yak_count(tmp_path):
    """The directly synthesised tables hold exactly the k-mer set `yak count` finds in error-free reads."""
    import subprocess
    G = synth.genome(5, 8000)
    fa = str(tmp_path / "r.fa")
    synth.write_short_reads(fa, 1, [G], depth_each=60.0, rlen=150, sub_rate=0.0)
    out = str(tmp_path / "k21.yak")
    subprocess.check_call([os.path.join(REF, "yak"), "count", "-k", "21", "-t", "2", "-o", out, fa], stderr=subprocess.DEVNULL)
    t = O.Table.load(out)
    h, c = synth.make_table(1, 21, [G], keep_min=1)
    got = t.lookup(h, 0)
    assert (got > 0).mean() > 0.999  # every genome k-mer was counted (up to uncovered ends)
    assert len(t) <= len(h)
