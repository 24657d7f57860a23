"""Developer aid: e2e phase breakdown with pinned host buffers."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import bench
import nextpolish2_b200 as np2
L = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
A, c, tabs = bench.make_workload(20260002, L, 16)
ctx = np2.Context(0)
tables = [np2.Table.from_arrays(ctx, k, *tabs[k]) for k in bench.KS]
bam = torch.from_numpy(c["bam"]).pin_memory().numpy()
ctg = torch.from_numpy(A.copy()).pin_memory().numpy()
for i in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); j = np2.Job(ctx, ctg, bam, tables, np2.Opts())
    t1 = time.perf_counter(); j.upload()
    t2 = time.perf_counter(); j.run(-1)
    t3 = time.perf_counter(); f, l, b = j.bases()
    t4 = time.perf_counter(); j.destroy()
    t5 = time.perf_counter()
    print("create %.1f upload %.1f run %.1f fetch %.1f destroy %.1f total %.1f ms" % tuple(1e3 * x for x in (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t5 - t0)))
