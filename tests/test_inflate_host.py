"""The DEFLATE decoder of the device inflate kernel (csrc/np2_inflate.cuh), compiled for the host, against zlib: every
block type, strategy and level zlib can emit, multi-block members, BGZF files written by the synthetic-input tool, and
corrupted payloads (must be rejected or agree with zlib, never crash).  No GPU needed."""
import ctypes as C
import os
import random
import subprocess
import zlib

import numpy as np
import pytest

import bgzf as OB  # oracle/bgzf.py
from nextpolish2_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("infl") / "libinflhost.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "nextpolish2_b200", "csrc"),
                           "-o", so, os.path.join(ROOT, "tests", "inflate_host_harness.cpp")])
    lib = C.CDLL(so)

    def infl(payload, cap):
        out = (C.c_uint8 * max(cap, 1))()
        n = C.c_uint32(0)
        ok = lib.np2t_inflate(bytes(payload), len(payload), out, cap, C.byref(n))
        assert ok >= 0, "result depends on the payload's alignment"
        return bool(ok), bytes(out[:n.value])
    return infl


def payloads(rng, n):
    kinds = [
        lambda: bytes(rng.getrandbits(8) for _ in range(n)),                      # incompressible
        lambda: bytes(rng.choice(b"ACGT") for _ in range(n)),
        lambda: b"\xff" * n,                                                        # QUAL of a BAM without qualities
        lambda: bytes(rng.choice([0x11, 0x12, 0x14, 0x18, 0x21, 0x22, 0x24, 0x28, 0x41, 0x42, 0x44, 0x48, 0x81, 0x82, 0x84,
                                  0x88]) for _ in range(n)),                        # 4-bit SEQ
        lambda: (b"ab" * n)[:n],                                                    # distance 2, overlapping matches
    ]

    def words():
        w = [bytes(rng.getrandbits(8) for _ in range(rng.randint(1, 40))) for _ in range(20)]
        b = b""
        while len(b) < n:
            b += rng.choice(w)
        return b[:n]
    return [k() for k in kinds] + [words()]


def test_every_block_type_matches_zlib(harness):
    rng = random.Random(1)
    n_streams = 0
    for n in (0, 1, 2, 3, 10, 100, 1000, 5000, 65280, 65536):
        for data in payloads(rng, n):
            for level in (0, 1, 6, 9):
                for strat in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE):
                    co = zlib.compressobj(level, zlib.DEFLATED, -15, rng.choice([1, 8, 9]), strat)
                    if n > 100 and rng.random() < 0.5:  # several blocks, with an empty stored block in between
                        k = rng.randint(1, n - 1)
                        p = co.compress(data[:k]) + co.flush(zlib.Z_SYNC_FLUSH) + co.compress(data[k:]) + co.flush()
                    else:
                        p = co.compress(data) + co.flush()
                    ok, o = harness(p, len(data))
                    assert ok and o == data, (n, level, strat)
                    n_streams += 1
                    if n:
                        assert not harness(p, len(data) - 1)[0]          # output larger than ISIZE
                    if len(p) > 8:
                        ok, o = harness(p[:len(p) // 2], len(data))      # truncated payload
                        assert not ok
    assert n_streams == 10 * 6 * 4 * 4


def test_bgzf_file_of_records(harness, tmp_path):
    A = synth.genome(11, 60_000)
    c = synth.make_contig(12, A, depth=20, asm_err=1e-4, het=0.0, mean_len=6000, sd_len=1000, min_len=2000, threads=2)
    for level in (0, 1, 6):
        path = str(tmp_path / ("t%d.bam" % level))
        synth.write_bam(path, ["ctg"], [len(A)], [c["bam"]], level=level)
        buf = open(path, "rb").read()
        ms = OB.members(buf)
        assert len(ms) > 3
        got = []
        for po, pl, isize, crc in ms:
            ok, o = harness(buf[po:po + pl], isize)
            assert ok and len(o) == isize and zlib.crc32(o) == crc
            got.append(o)
        assert b"".join(got) == OB.inflate_all(buf)
        assert bytes(c["bam"]) in b"".join(got)


def test_corrupt_payloads_never_crash_and_agree_with_zlib(harness):
    rng = random.Random(5)
    data = payloads(rng, 30000)[5]
    p = zlib.compress(data, 6)[2:-4]
    for it in range(1500):
        q = bytearray(p)
        for _ in range(rng.randint(1, 4)):
            q[rng.randrange(len(q))] ^= 1 << rng.randrange(8)
        ok, o = harness(bytes(q), len(data))
        try:
            d = zlib.decompressobj(-15)
            ref = d.decompress(bytes(q))
            ref = ref if d.eof else None
        except zlib.error:
            ref = None
        if ok:  # a complete valid stream of at most ISIZE bytes: zlib must read the same one
            assert ref == o
        elif ref is not None:  # zlib accepts: the only reason to refuse is an output beyond the member's ISIZE
            assert len(ref) > len(data)


def test_member_walk_of_the_ctypes_mirror_matches_the_oracle(tmp_path):
    """api.bgzf_members (what the tests and the bench hand to np2_bgzf_inflate / np2_job_create_bgzf) against the
    oracle's walk, including the empty EOF member; garbage is refused."""
    import nextpolish2_b200 as np2
    A = synth.genome(31, 50_000)
    c = synth.make_contig(32, A, depth=15, asm_err=1e-4, het=0.0, mean_len=5000, sd_len=800, min_len=2000, threads=2)
    path = str(tmp_path / "m.bam")
    synth.write_bam(path, ["ctg"], [len(A)], [c["bam"]], level=6)
    buf = np.fromfile(path, np.uint8)
    po, pl, iz = np2.bgzf_members(buf)
    ms = OB.members(buf)
    assert [int(x) for x in po] == [m[0] for m in ms]
    assert [int(x) for x in pl] == [m[1] for m in ms]
    assert [int(x) for x in iz] == [m[2] for m in ms]
    assert iz[-1] == 0 and pl[-1] == 2  # the BGZF end-of-file marker: an empty fixed-Huffman block
    with pytest.raises(ValueError):
        np2.bgzf_members(buf[5:])
    with pytest.raises(ValueError):
        np2.bgzf_members(buf[:len(buf) - 9])
