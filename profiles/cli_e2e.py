"""End-to-end timing of the nextPolish2 command line on the GPU box: process start -> FASTA written, BAM (BGZF) decode
and table loading included (SURVEY 8(d) metric (1)).   usage: python profiles/cli_e2e.py [n_contigs] [contig_bp] [bgzf levels, e.g. 1,0]"""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nextpolish2_b200 import synth  # noqa: E402

n_ctg = int(sys.argv[1]) if len(sys.argv) > 1 else 4
L = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
cli = os.path.join(ROOT, "nextpolish2_b200", "nextPolish2")
with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
    names, contigs, blobs, haps = [], [], [], []
    G = synth.genome(20260002, n_ctg * L)
    t0 = time.time()
    for i in range(n_ctg):
        A = G[i * L:(i + 1) * L].copy()
        c = synth.make_contig(20260100 + i, A, ref_id=i, depth=30.0, asm_err=2e-5, het=0.0, read_err=0.002, threads=16)
        names.append("ctg%d" % i)
        contigs.append(A)
        blobs.append(c["bam"])
        haps.append(c["hap1"])
    yaks = []
    for k in (21, 31):
        p = os.path.join(d, "k%d.yak" % k)
        synth.write_yak(p, k, *synth.make_table(20260003, k, haps))
        yaks.append(p)
    fa = os.path.join(d, "asm.fa")
    synth.write_fasta(fa, names, contigs, width=80)
    print("inputs generated in %.1f s" % (time.time() - t0), flush=True)
    for level in [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else '1,0').split(',')]:
        bam = os.path.join(d, "hifi.l%d.bam" % level)
        t0 = time.time()
        synth.write_bam(bam, names, [L] * n_ctg, blobs, level=level)
        print("BAM level %d: %.1f MB written in %.1f s" % (level, os.path.getsize(bam) / 1e6, time.time() - t0), flush=True)
        for rep in range(2):
            out = os.path.join(d, "out.fa")
            if os.path.exists(out):
                os.remove(out)
            t0 = time.time()
            r = subprocess.run([cli, "-t", "16", "-o", out, bam, fa] + yaks, capture_output=True, text=True,
                               env=dict(os.environ, NP2_CLI_TIMING="1"))
            dt = time.time() - t0
            ok = r.returncode == 0 and all(bytes(h) in open(out, "rb").read() for h in haps[:1])
            print("level %d run %d: %.2f s wall (%.1f Mbp/s), rc %d, first contig == truth haplotype: %s\n   %s" % (
                level, rep, dt, n_ctg * L / 1e6 / dt, r.returncode, ok, r.stderr.strip().splitlines()[-1] if r.stderr.strip() else ""),
                flush=True)
        os.remove(bam)
