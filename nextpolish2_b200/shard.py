"""Contig-level sharding across the GPUs of one box.

Contigs are independent (the reference's worker closure shares nothing but read-only inputs, src/main.rs:1726-1838),
so ranks get a longest-processing-time-first share of the contigs, every rank keeps a full replica of the yak tables
and there is NO collective on the data path; the only communication is gathering the finished FASTA records so that
rank 0 can emit them in input order (= the reference with -t 1; with -t > 1 its order is arbitrary, main.rs:1845-1851).
"""
import heapq


def lpt_partition(weights, n_parts):
    """Longest-processing-time-first: returns n_parts lists of item indices (each list in input order)."""
    parts = [[] for _ in range(n_parts)]
    heap = [(0.0, p) for p in range(n_parts)]
    heapq.heapify(heap)
    for i in sorted(range(len(weights)), key=lambda i: (-weights[i], i)):
        load, p = heapq.heappop(heap)
        parts[p].append(i)
        heapq.heappush(heap, (load + weights[i], p))
    return [sorted(p) for p in parts]


def polish_sharded(n_items, weights, polish_fn, rank=0, world=1, group=None):
    """Runs polish_fn(i) -> bytes for this rank's share; returns the records in input order on rank 0 (None elsewhere).

    `weights[i]` ~ contig length x depth.  With world == 1 no process group is needed."""
    mine = lpt_partition(weights, world)[rank]
    local = {i: polish_fn(i) for i in mine}
    if world == 1:
        return [local[i] for i in range(n_items)]
    import torch.distributed as dist
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(local, gathered, dst=0, group=group)
    if rank != 0:
        return None
    merged = {}
    for d in gathered:
        merged.update(d)
    return [merged[i] for i in range(n_items)]
