"""Randomised campaign (not collected by pytest): the C++ oracle against the pure-Python restatement on the whole path,
and the library's host phasing against the oracle on random read graphs.   usage: python tests/fuzz_restatements.py
[seconds per campaign] [seed].  CPU only.  Every case draws contig length, depth, error / het rates, read lengths, clips,
supplementary / low-MAPQ reads, N and lower-case stretches in the draft, =/X or M CIGARs, one to three k-mer tables
(k < 32 and k >= 32) and the options -m / -r / -i / -n / -c; a mismatch prints the case and the traceback."""
import os
import sys
import time
import traceback

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(HERE), os.path.join(os.path.dirname(HERE), "oracle"), HERE]
import oracle as O  # noqa: E402
import test_phase_host as TP  # noqa: E402
import test_py_restatement as T  # noqa: E402
from nextpolish2_b200 import api, synth  # noqa: E402


def full_path(seconds, rng):
    t0, n, bad = time.time(), 0, 0
    while time.time() - t0 < seconds:
        seed = int(rng.integers(1 << 30))
        L = int(rng.integers(4000, 9000))
        A = synth.genome(seed, L).copy()
        if rng.random() < 0.3:
            p = int(rng.integers(100, L - 200))
            A[p:p + int(rng.integers(1, 30))] = ord("N")
        if rng.random() < 0.3:
            p = int(rng.integers(100, L - 200))
            A[p:p + 50] |= 0x20
        kw = dict(depth=int(rng.integers(8, 45)), asm_err=float(rng.choice([1e-4, 1e-3, 5e-3])),
                  het=float(rng.choice([0, 0.002, 0.01, 0.02])), mean_len=int(rng.integers(1500, 4000)), sd_len=400,
                  min_len=1100, read_err=float(rng.choice([0.001, 0.006, 0.02])), frac_clip=float(rng.choice([0, 0.1])),
                  frac_lowq=0.03, frac_supp=0.05, eqx=bool(rng.random() < 0.5), threads=2)
        ks = [21, 31] if rng.random() < 0.6 else ([21] if rng.random() < 0.5 else [25, 51])
        opt = {}
        r = rng.random()
        if r < 0.2:
            opt["model"] = 1
        elif r < 0.4:
            opt["use_all_reads"] = 1
        elif r < 0.5:
            opt["iter_count"] = 3
        elif r < 0.6:
            opt["iter_count"] = 1
        if rng.random() < 0.2:
            opt["max_indel_len"] = int(rng.integers(0, 5))
        if rng.random() < 0.2:
            opt["min_kmer_count"] = int(rng.integers(1, 30))
        try:
            c = synth.make_contig(seed + 1, A, **kw)
            tabs = {k: synth.make_table(seed + 2, k, [c["hap1"], c["hap2"]], mean_count=float(rng.choice([8, 40]))) for k in ks}
            T.compare_full(A, c["bam"], tabs, **opt)
            n += 1
        except Exception:
            bad += 1
            print("MISMATCH full path: seed", seed, L, kw, ks, opt)
            traceback.print_exc(limit=3)
            sys.stdout.flush()
    print("full path: %d cases agree, %d do not" % (n, bad))
    return bad


def phasing(seconds, rng):
    t0, n, bad, paths = time.time(), 0, 0, {}
    while time.time() - t0 < seconds:
        kw = dict(seed=int(rng.integers(1 << 30)), n_reads=int(rng.integers(5, 600)), span=int(rng.integers(2, 60)),
                  noise=float(rng.choice([0, 0.02, 0.1, 0.3, 0.5])), gap_every=int(rng.choice([0, 0, 20, 100])),
                  ref=bool(rng.random() < 0.8))
        keys, vals = TP.make_graph(**kw)
        if rng.random() < 0.3:  # a community that falls apart again (louvain.rs:136-165)
            gk, gv = TP.gadget(base=kw["n_reads"] + int(rng.integers(1, 50)))
            keys, vals = np.concatenate([keys, gk]), np.concatenate([vals, gv])
            o = np.argsort(keys, kind="stable")
            keys, vals = keys[o], vals[o]
        for model, use_all in ((0, False), (1, False), (0, True), (1, True)):
            got, path = api.debug_phase(keys, vals, model, use_all, with_path=True)
            paths[path] = paths.get(path, 0) + 1
            n += 1
            if not np.array_equal(np.sort(got), O.debug_phase(keys, vals, model, use_all)):
                bad += 1
                print("MISMATCH phasing:", kw, model, use_all)
    print("phasing: %d cases agree, %d do not (flat-array path %d, general path %d)" % (n - bad, bad, paths.get(1, 0), paths.get(2, 0)))
    return bad


if __name__ == "__main__":
    secs = float(sys.argv[1]) if len(sys.argv) > 1 else 300
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 777)
    sys.exit(1 if full_path(secs, rng) + phasing(secs, rng) else 0)
