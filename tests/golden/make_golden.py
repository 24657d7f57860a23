"""Generates the committed golden fixtures from the COMPILED REFERENCE C code (oracle/_ref, built by
oracle/Makefile from /root/reference/yak).  Run in the build container only:

    python tests/golden/make_golden.py

Outputs (small, committed):
    k21.yak k31.yak k51.yak   tables written by the reference `yak count` on a seeded 3 kb genome's reads
    yak_kat.json              hash known answers (yak-priv.h:10-38 via libyakref.so) and k-mer -> count
                              answers from the reference's own yak_ch_restore + yak_ch_get (htab.c:80,213)
"""
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from nextpolish2_b200 import synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")


def main():
    lib = C.CDLL(os.path.join(REF, "libyakref.so"))
    lib.yakref_hash64.restype = C.c_uint64
    lib.yakref_hash64.argtypes = [C.c_uint64, C.c_uint64]
    lib.yakref_hash64_64.restype = C.c_uint64
    lib.yakref_hash64_64.argtypes = [C.c_uint64]
    lib.yakref_hash_long.restype = C.c_uint64
    lib.yakref_hash_long.argtypes = [C.c_void_p]
    lib.yakref_restore.restype = C.c_void_p
    lib.yakref_restore.argtypes = [C.c_char_p]
    lib.yakref_get.argtypes = [C.c_void_p, C.c_uint64]
    rng = np.random.default_rng(20260001)
    kat = {"hash64": [], "hash64_64": [], "hash_long": [], "tables": {}}
    for _ in range(64):
        k = int(rng.integers(1, 32))
        mask = (1 << (2 * k)) - 1
        key = int(rng.integers(0, 2**63)) & mask
        kat["hash64"].append([key, mask, lib.yakref_hash64(key, mask)])
    for _ in range(32):
        key = int(rng.integers(0, 2**63)) * 2 + int(rng.integers(0, 2))
        kat["hash64_64"].append([key, lib.yakref_hash64_64(key)])
    for _ in range(32):
        x = (C.c_uint64 * 4)(*[int(v) for v in rng.integers(0, 2**62, 4)])
        kat["hash_long"].append([list(x), lib.yakref_hash_long(x)])

    G = synth.genome(20260001, 3000)
    with tempfile.TemporaryDirectory() as td:
        fa = os.path.join(td, "sr.fa")
        synth.write_short_reads(fa, 7, [G], depth_each=30.0, rlen=150, sub_rate=0.002)
        for k in (21, 31, 51):
            out = os.path.join(HERE, "k%d.yak" % k)
            subprocess.check_call([os.path.join(REF, "yak"), "count", "-k", str(k), "-t", "1", "-o", out, fa],
                                  stderr=subprocess.DEVNULL)
            h = lib.yakref_restore(out.encode())
            g = bytes(G)
            kmers, counts = [], []
            for i in range(48):  # present k-mers, both strands
                s = int(rng.integers(0, len(g) - k))
                kmers.append(g[s:s + k].decode())
            comp = bytes.maketrans(b"ACGT", b"TGCA")
            for i in range(8):
                s = int(rng.integers(0, len(g) - k))
                kmers.append(g[s:s + k].translate(comp)[::-1].decode())
            for i in range(16):  # random (absent) k-mers
                kmers.append("".join("ACGT"[int(x)] for x in rng.integers(0, 4, k)))
            # the reference's own hashing of these strings is checked in test_oracle_yak (CPU): here we ask the
            # reference table with the oracle-independent python hash below
            for s in kmers:
                counts.append(lib.yakref_get(h, py_hash(s, k, lib)))
            n_keys = (os.path.getsize(out) - 16 - 1024 * 8) // 8
            kat["tables"][str(k)] = {"n_keys": n_keys, "kmers": kmers, "counts": counts}
    json.dump(kat, open(os.path.join(HERE, "yak_kat.json"), "w"), indent=0)
    print("wrote", os.path.join(HERE, "yak_kat.json"))


def py_hash(s, k, lib):
    """canonical k-mer hash exactly as yak/count.c:28-74 computes it, using the reference's hash functions"""
    code = {"A": 0, "C": 1, "G": 2, "T": 3}
    if k < 32:
        mask = (1 << (2 * k)) - 1
        f = r = 0
        for ch in s:
            c = code[ch]
            f = (f << 2 | c) & mask
            r = r >> 2 | (3 - c) << (2 * (k - 1))
        return lib.yakref_hash64(min(f, r), mask)
    mask = (1 << k) - 1
    x = [0, 0, 0, 0]
    for ch in s:
        c = code[ch]
        x[0] = (x[0] << 1 | (c & 1)) & mask
        x[1] = (x[1] << 1 | (c >> 1)) & mask
        x[2] = x[2] >> 1 | (1 - (c & 1)) << (k - 1)
        x[3] = x[3] >> 1 | (1 - (c >> 1)) << (k - 1)
    arr = (C.c_uint64 * 4)(*x)
    return lib.yakref_hash_long(arr)


if __name__ == "__main__":
    main()
