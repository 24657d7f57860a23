"""One 10 Mbp contig through np2_job_create_bgzf -> upload -> run (twice: the second pass runs with the remembered
capacities), for the ncu launch list of the file -> FASTA path:
    ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file out.csv python profiles/bgzf_job_launches.py"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nextpolish2_b200 as np2  # noqa: E402
from nextpolish2_b200 import synth  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
A = synth.genome(20260002, L)
c = synth.make_contig(20260100, A, ref_id=0, depth=30.0, asm_err=2e-5, het=0.0, read_err=0.002, threads=16)
tabs = {k: synth.make_table_mt(20260003, k, [c["hap1"]], threads=16) for k in (21, 31)}
with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
    path = os.path.join(d, "c.bam")
    synth.write_bam(path, ["ctg0"], [L], [c["bam"]], level=1)
    buf = np.fromfile(path, np.uint8)
ctx = np2.Context(0)
gt = [np2.Table.from_arrays(ctx, k, *tabs[k]) for k in (21, 31)]
po, pl, iz = np2.bgzf_members(buf)
rec = bytes(c["bam"])
inflated, _ = np2.bgzf_inflate(ctx, buf, po, pl, iz)
skip = bytes(inflated).find(rec[:4096])
del inflated
for rep in range(2):
    job = np2.Job.from_bgzf(ctx, A, buf, po, pl, iz, skip, len(rec), gt, np2.Opts()).upload().run()
    first, last, base = job.bases()
    assert bytes(base) == bytes(c["hap1"]), "polished contig differs from the truth haplotype"
    job.destroy()
print("ok: %d members -> %d bp polished, identical to the truth haplotype" % (len(po), L))
