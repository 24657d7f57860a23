"""CPU: the multi-GPU path's host logic with world_size = 2 over gloo.  The per-contig work is stubbed with the oracle
(tests may use it); what is under test is the partition, the absence of any data-path collective and that rank 0 gets
the records back in input order, byte-identical to a single-process run."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lpt_partition_properties():
    from nextpolish2_b200.shard import lpt_partition
    rng = np.random.default_rng(0)
    for n, k in [(24, 8), (5, 8), (1, 2), (100, 3)]:
        w = rng.integers(1, 250, n).tolist()
        parts = lpt_partition(w, k)
        assert sorted(i for p in parts for i in p) == list(range(n))
        loads = [sum(w[i] for i in p) for p in parts]
        assert max(loads) <= sum(w) / k + max(w)  # LPT bound
    assert lpt_partition([5, 5, 5, 5], 2) == [[0, 2], [1, 3]]


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import torch.distributed as dist
    import oracle as O
    from nextpolish2_b200 import synth
    from nextpolish2_b200.shard import polish_sharded
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    lens = [30_000, 12_000, 45_000, 20_000, 8_000]
    G = synth.genome(77, sum(lens))
    contigs, off = [], 0
    for i, L in enumerate(lens):
        A = G[off:off + L].copy()
        off += L
        contigs.append((A, synth.make_contig(100 + i, A, depth=15, asm_err=5e-4, mean_len=4000, sd_len=500, min_len=2000, threads=1)))
    tabs = [O.Table.from_arrays(k, *synth.make_table(9, k, [c[1]["hap1"] for c in contigs])) for k in (21, 31)]

    def polish(i):
        j = O.Job(contigs[i][0], contigs[i][1]["bam"], tabs, O.Opts(min_ctg_len=0))
        return O.format_fasta("ctg%d" % i, *j.consensus())
    recs = polish_sharded(len(lens), [float(x) for x in lens], polish, rank, world)
    if rank == 0:
        single = [polish(i) for i in range(len(lens))]
        q.put(recs == single and all(r.startswith(b">ctg%d " % i) for i, r in enumerate(recs)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_gloo_input_order():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    ok = q.get(timeout=240)
    [p.join(60) for p in procs]
    assert ok and all(p.exitcode == 0 for p in procs)
