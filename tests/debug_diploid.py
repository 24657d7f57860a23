"""Developer aid: stage times on a diploid contig (phasing path)."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import numpy as np
import nextpolish2_b200 as np2
from nextpolish2_b200 import synth
L = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
t0 = time.time()
A = synth.genome(31, L)
c = synth.make_contig(32, A, depth=30.0, asm_err=2e-5, het=0.01, read_err=0.002, threads=16)
tabs = {k: synth.make_table(33, k, [c["hap1"], c["hap2"]]) for k in (21, 31)}
print("gen %.1fs reads %d" % (time.time() - t0, c["n_reads"]), flush=True)
ctx = np2.Context(0)
tables = [np2.Table.from_arrays(ctx, k, *tabs[k]) for k in (21, 31)]
job = np2.Job(ctx, A, c["bam"], tables, np2.Opts()).upload()
for i in range(3):
    t = time.time(); job.run(-1); print("run %.1f ms" % ((time.time() - t) * 1e3), flush=True)
f, l, b = job.bases()
print("identical to hap1:", bytes(b) == bytes(c["hap1"]), len(b), len(c["hap1"]))
tm = job.timings()
for k, v in sorted(tm.items(), key=lambda x: -x[1][0])[:22]:
    print("%-28s %9.3f ms" % (k, v[0]))
print(job.traffic())
if len(sys.argv) > 2:
    import oracle as O
    oj = O.Job(A, c["bam"], [O.Table.from_arrays(k, *tabs[k]) for k in (21, 31)], O.Opts())
    print("oracle %.1fs identical:" % oj.seconds, bytes(oj.consensus()[1]) == bytes(b))
