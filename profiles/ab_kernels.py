"""A/B of kernel variants inside one process on the bench workload (10 Mbp haploid, 30x): the launch wrappers read
NP2_PILE_BATCH / NP2_PACK_BATCH on every call, so the same resident job is stepped under each setting.  Prints the
library's own CUDA-event timers for the two kernels and the whole step, and checks that the consensus is byte-identical
under every setting.   usage (GPU box): python profiles/ab_kernels.py [steps]"""
import hashlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    import torch
    import nextpolish2_b200 as np2
    A, c, tabs = bench.make_workload(20260002, 10_000_000, 16, 0.0)
    ctx = np2.Context(0)
    tables = [np2.Table.from_arrays(ctx, k, *tabs[k]) for k in bench.KS]
    bam = torch.from_numpy(c["bam"]).pin_memory().numpy()
    contig = torch.from_numpy(A.copy()).pin_memory().numpy()
    job = np2.Job(ctx, contig, bam, tables, np2.Opts()).upload()
    variants = [("old", 0, 1), ("pile2", 2, 1), ("pile4", 4, 1), ("pack2", 0, 2), ("pile4+pack2", 4, 2),
                ("pile2+pack2", 2, 2), ("old again", 0, 1)]
    digests = {}
    for name, pile, pack in variants:
        os.environ["NP2_PILE_BATCH"] = str(pile)
        os.environ["NP2_PACK_BATCH"] = str(pack)
        for _ in range(4):
            job.run(-1)
        acc = {}
        for _ in range(steps):
            job.run(-1)
            for k, v in job.timings().items():
                acc.setdefault(k, []).append(v[0])
        pos, base = job.consensus()
        digests[name] = hashlib.sha256(bytes(base) + np.asarray(pos).tobytes()).hexdigest()[:16]
        print(json.dumps({"variant": name, "pileup_emit_ms": round(float(np.mean(acc["pileup_emit"])), 4),
                          "pack_columns_ms": round(float(np.mean(acc["pack_columns"])), 4),
                          "step_ms": round(float(np.mean(acc["total"])), 4),
                          "step_ms_min": round(float(np.min(acc["total"])), 4),
                          "truth": bytes(base) == bytes(c["hap1"]), "digest": digests[name]}), flush=True)
    assert len(set(digests.values())) == 1, digests
    print("all variants byte-identical")


if __name__ == "__main__":
    main()
