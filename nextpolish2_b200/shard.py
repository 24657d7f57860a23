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


class _DeviceBytes:
    """A device pointer as a __cuda_array_interface__ object, so that torch can wrap it without a copy."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def table_meta(table):
    """What the receiving ranks need besides the image bytes."""
    ptr, nbytes, nb = table.image()
    return {"k": table.k, "n_keys": len(table), "buckets_per_subtable": nb, "bytes": nbytes}


def broadcast_tables(ctx, tables, rank=0, world=1, src=0, group=None):
    """One rank stages the yak tables (from .yak files or arrays); every other rank of the box receives the device
    images over NVLink (ncclBroadcast through torch.distributed) and adopts them, instead of each rank reading and
    staging its own copy over PCIe.  `tables` is the list of staged Tables on `src` and ignored elsewhere.
    Returns this rank's tables (the originals on `src`)."""
    if world == 1:
        return list(tables)
    import torch
    import torch.distributed as dist
    from .api import Table
    metas = [[table_meta(t) for t in tables] if rank == src else None]
    dist.broadcast_object_list(metas, src=src, group=group)
    out = []
    for i, m in enumerate(metas[0]):
        if rank == src:
            ptr, nbytes, _ = tables[i].image()
            buf = torch.as_tensor(_DeviceBytes(ptr, nbytes), device="cuda")  # the staged image itself, no copy
            dist.broadcast(buf, src=src, group=group)
            out.append(tables[i])
        else:
            buf = torch.empty(m["bytes"], dtype=torch.uint8, device="cuda")
            dist.broadcast(buf, src=src, group=group)
            torch.cuda.synchronize()
            out.append(Table.adopt(ctx, m["k"], m["n_keys"], m["buckets_per_subtable"], buf.data_ptr(), m["bytes"]))
            del buf
    return out
