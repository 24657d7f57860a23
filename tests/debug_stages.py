"""Developer aid (not a test): prints every stage mismatch between oracle and GPU for one dataset."""
import sys
import traceback

sys.path.insert(0, ".")
sys.path.insert(0, "oracle")
sys.path.insert(0, "tests")
import numpy as np

import common
import oracle as O
import nextpolish2_b200 as np2


def cmp(name, a, b):
    try:
        common.assert_same(name, a, b)
        print("  ok   ", name, len(a), flush=True)
        return True
    except AssertionError as e:
        print("  DIFF ", str(e)[:400], flush=True)
        return False


def main():
    name = sys.argv[1]
    it = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    kw = {}
    for a in sys.argv[3:]:
        k, v = a.split("=")
        kw[k] = int(v)
    ds = common.dataset(name)
    oo, go = common.same_opts(**kw)
    ctx = np2.Context(0)
    print("dataset", name, "reads", ds["n_reads"], "iter", it, flush=True)
    oj = O.Job(ds["contig"], ds["bam"], common.oracle_tables(ds), oo, dump_iter=it)
    print("oracle done", flush=True)
    gj = np2.Job(ctx, ds["contig"], ds["bam"], common.gpu_tables(ctx, ds), go)
    try:
        gj.upload().run(it)
        print("gpu done", gj.timings(), flush=True)
    except Exception:
        traceback.print_exc()
    for stage, keys in [("reads", ["rec_idx", "t_s", "t_e", "blank", "nib_off", "nib"]),
                        ("msa", ["off", "bases", "delta", "count", "besti"]),
                        ("dp_consensus", ["pos", "base", "flags"]),
                        ("regions", ["start", "end", "lable"]),
                        ("candidates", ["roff", "order", "seq_off", "seq", "kmer", "kscore"])]:
        try:
            a, b = getattr(oj, stage)(), getattr(gj, stage)()
            for k in keys:
                cmp(stage + "." + k, a[k], b[k])
        except Exception:
            traceback.print_exc()
    cmp("dropped", oj.dropped(), gj.dropped())
    op, ob = oj.consensus()
    gp, gb = gj.consensus()
    cmp("final.base", ob, gb)
    cmp("final.pos", op, gp)


main()
