"""Developer aid: one 10 Mbp job, print every stage time."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import bench
import nextpolish2_b200 as np2
L = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
A, c, tabs = bench.make_workload(20260002, L, 16)
ctx = np2.Context(0)
tables = [np2.Table.from_arrays(ctx, k, *tabs[k]) for k in bench.KS]
t0 = time.time(); job = np2.Job(ctx, A, c["bam"], tables, np2.Opts()); t1 = time.time()
job.upload(); t2 = time.time()
print("create %.1f ms upload %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
for i in range(3):
    t = time.time(); job.run(-1); print("run %.1f ms" % ((time.time() - t) * 1e3))
tm = job.timings()
for k, v in sorted(tm.items(), key=lambda x: -x[1][0]):
    print("%-28s %9.3f ms  launches %d" % (k, v[0], v[1]))
print(job.traffic())
