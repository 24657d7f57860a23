"""End-to-end timing of the nextPolish2 command line on the GPU box: process start -> FASTA written, BAM (BGZF) decode
and table loading included (SURVEY 8(d) metric (1)).
usage: python profiles/cli_e2e.py [n_contigs] [contig_bp] [bgzf levels, e.g. 1,0] [n_distinct] [timing level] [lanes]
n_distinct < n_contigs: only that many contigs are synthesised, the others are copies under their own name / refID (the
generator, not the command line, is what takes the time on the box)."""
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

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
    n_distinct = min(n_ctg, int(sys.argv[4])) if len(sys.argv) > 4 else n_ctg
    timing_level = sys.argv[5] if len(sys.argv) > 5 else "1"
    for i in range(n_ctg):
        names.append("ctg%d" % i)
        if i < n_distinct:
            A = G[i * L:(i + 1) * L].copy()
            c = synth.make_contig(20260100 + i, A, ref_id=i, depth=30.0, asm_err=2e-5, het=0.0, read_err=0.002, threads=16)
            contigs.append(A)
            blobs.append(c["bam"])
            haps.append(c["hap1"])
        else:  # a copy of contig i % n_distinct with its own refID in every record
            src = i % n_distinct
            b = np.array(blobs[src], np.uint8, copy=True)
            o, n = 0, len(b)
            rid = np.frombuffer(np.array([i], "<i4").tobytes(), np.uint8)
            while o < n:
                bs = int(b[o]) | int(b[o + 1]) << 8 | int(b[o + 2]) << 16 | int(b[o + 3]) << 24
                b[o + 4:o + 8] = rid
                o += 4 + bs
            contigs.append(contigs[src])
            blobs.append(b)
            haps.append(haps[src])
    yaks = []
    for k in (21, 31):
        p = os.path.join(d, "k%d.yak" % k)
        synth.write_yak(p, k, *synth.make_table(20260003, k, haps[:n_distinct]))
        yaks.append(p)
    fa = os.path.join(d, "asm.fa")
    synth.write_fasta(fa, names, contigs, width=80)
    print("inputs generated in %.1f s" % (time.time() - t0), flush=True)
    for level in [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else '1,0').split(',')]:
        bam = os.path.join(d, "hifi.l%d.bam" % level)
        t0 = time.time()
        synth.write_bam(bam, names, [L] * n_ctg, blobs, level=level)
        print("BAM level %d: %.1f MB written in %.1f s" % (level, os.path.getsize(bam) / 1e6, time.time() - t0), flush=True)
        modes = [("device inflate, records stay on the device (default)", [], {}),
                 ("device inflate, records come back to the host", [], {"NP2_CLI_RECORDS_ON_HOST": "1"}),
                 ("host inflate (zlib)", ["--host-inflate"], {})]
        if len(sys.argv) > 6 and sys.argv[6] == "lanes":  # contigs in flight per GPU on the default path
            modes = [("records stay on the device, %d lanes" % k, [], {"NP2_CLI_LANES": str(k)}) for k in (3, 4, 5, 6)]
        for label, mode, env in modes:
            for rep in range(1 if len(sys.argv) > 6 else 2):
                out = os.path.join(d, "out.fa")
                if os.path.exists(out):
                    os.remove(out)
                t0 = time.time()
                r = subprocess.run([cli, "-t", "16", "-o", out] + mode + [bam, fa] + yaks, capture_output=True, text=True,
                                   env=dict(os.environ, NP2_CLI_TIMING=timing_level, **env))
                dt = time.time() - t0
                got = open(out, "rb").read() if r.returncode == 0 else b""
                ok = r.returncode == 0 and all(bytes(h) in got for h in haps)
                lines = [l for l in r.stderr.strip().splitlines() if l.startswith("[np2")]
                done = sorted(float(l.split("done at ")[1].split(" s")[0]) for l in lines if l.startswith("[np2 contig]"))
                skip = 2 * int(env.get("NP2_CLI_LANES", 3))  # every lane's first contig pays the first-use costs
                steady = ""
                if len(done) > skip + 4:
                    per = (done[-1] - done[skip - 1]) / (len(done) - skip)
                    steady = "; steady state after the first %d contigs: %.1f ms per contig = %.0f Mbp/s" % (skip, per * 1e3, L / 1e6 / per)
                print("level %d %s run %d: %.2f s wall (%.1f Mbp/s)%s, rc %d, every contig == its truth haplotype: %s\n   %s" % (
                    level, label, rep, dt, n_ctg * L / 1e6 / dt, steady, r.returncode, ok, "\n   ".join(lines)), flush=True)
        os.remove(bam)
