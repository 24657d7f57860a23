"""How often can the reference's result depend on FxHashMap iteration order?  (SURVEY hard part 3)

The Rust reference iterates FxHashMaps at louvain.rs:123,145-165,199; the oracle and the product iterate in ascending
id order.  The order can only change a result in two places, which the oracle counts (np2o_order_exposure):
communities declustered in second_stage, and conflicting communities with EQUAL sort keys in phase_communities.
This script runs the oracle over real and synthetic diploid inputs and writes the table DESIGN.md quotes.

    python tests/order_exposure.py [--big]      # --big adds the 10 Mbp diploid contig of configs[2]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402

import oracle as O  # noqa: E402
from nextpolish2_b200 import synth  # noqa: E402


def run(name, contig, bam, tables, **okw):
    O.order_exposure(reset=True)
    t0 = time.time()
    rows = {}
    for label, kw in (("ref model", {}), ("len model (-m len)", {"model": 1}), ("-r (use all reads)", {"use_all_reads": 1})):
        kw = dict(kw, **okw)
        j = O.Job(contig, bam, tables, O.Opts(**kw), dump_iter=-1)
        e = O.order_exposure(reset=True)
        e["reads_dropped"] = int(len(j.dropped()))
        rows[label] = e
    print("%-28s %5.1fs  %s" % (name, time.time() - t0, json.dumps(rows)), flush=True)
    return rows


def main():
    import common
    import gzip
    out = {}
    gold = os.path.join(ROOT, "tests", "golden")
    for fx in ("c1_40k", "_c1"):
        d = os.path.join(gold, fx)
        if not os.path.exists(os.path.join(d, "k21.yak")):
            continue
        if fx == "c1_40k":
            contig = np.frombuffer(gzip.open(os.path.join(d, "contig.bin.gz")).read(), np.uint8)
            bam = np.frombuffer(gzip.open(os.path.join(d, "records.bin.gz")).read(), np.uint8)
        else:
            contig = np.fromfile(os.path.join(d, "contig.bin"), np.uint8)
            bam = np.fromfile(os.path.join(d, "records.bin"), np.uint8)
        tabs = [O.Table.load(os.path.join(d, "k%d.yak" % k)) for k in (21, 31)]
        out["configs[0] real HiFi reads (%s, %d bp)" % (fx, len(contig))] = run(fx, contig, bam, tabs, min_ctg_len=0)
    for name in ("dip600k", "clip120k", "tandem200k", "deep80k"):
        ds = common.dataset(name)
        out["synthetic " + name] = run(name, ds["contig"], ds["bam"], common.oracle_tables(ds), min_ctg_len=0)
    if "--big" in sys.argv:
        L = 10_000_000
        A = synth.genome(20260002, L)
        c = synth.make_contig(20260003, A, depth=30.0, asm_err=2e-5, het=0.01, read_err=0.002, threads=8)
        tabs = [O.Table.from_arrays(k, *synth.make_table_mt(20260004, k, [c["hap1"], c["hap2"]], threads=8)) for k in (21, 31)]
        out["synthetic 10 Mbp diploid, 1% het (one contig of configs[2])"] = run("dip10M", A, c["bam"], tabs)
    path = os.path.join(ROOT, "profiles", "r02_order_exposure.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
