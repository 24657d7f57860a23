"""BGZF inflate of one contig's records: np2_bgzf_inflate (device, one warp per member) against zlib on the host cores.
usage: python profiles/bgzf_inflate_bench.py [contig_bp] [levels, e.g. 1,6] [sweep]   (run on the GPU box)
sweep: also time the kernel for every combination of the tuning knobs of np2_inflate.cu."""
import os
import sys
import time
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nextpolish2_b200 as np2  # noqa: E402
from nextpolish2_b200 import synth  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
levels = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1,6").split(",")]
A = synth.genome(20260002, L)
c = synth.make_contig(20260100, A, ref_id=0, depth=30.0, asm_err=2e-5, het=0.0, read_err=0.002, threads=16)
ctx = np2.Context(0)
cores = os.cpu_count()
d = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
for level in levels:
    path = os.path.join(d, "np2_bgzf_bench.l%d.bam" % level)
    synth.write_bam(path, ["ctg0"], [L], [c["bam"]], level=level)
    buf = np.fromfile(path, np.uint8)
    os.remove(path)
    po, pl, iz = np2.bgzf_members(buf)
    total = int(iz.astype(np.uint64).sum())
    raw = bytes(buf)

    def host(i):
        return zlib.decompress(raw[int(po[i]):int(po[i]) + int(pl[i])], -15)
    with ThreadPoolExecutor(cores) as ex:
        list(ex.map(host, range(0, len(po), 8)))
        t0 = time.perf_counter()
        parts = list(ex.map(host, range(len(po)), chunksize=64))
        t_host = time.perf_counter() - t0
    want = b"".join(parts)
    pin_in, pin_out = np2.PinnedBuffer(buf), np2.PinnedBuffer(np.zeros(total, np.uint8))
    res = {}
    for name, src in (("pinned", pin_in), ("pageable", buf)):
        ks, ws = [], []
        for rep in range(6):
            t0 = time.perf_counter()
            got, kms = np2.bgzf_inflate(ctx, src, po, pl, iz, out=pin_out)
            ws.append((time.perf_counter() - t0) * 1e3)
            ks.append(kms)
        assert bytes(got) == want, "device inflate differs from zlib"
        res[name] = (min(ks[1:]), min(ws[1:]))
    if len(sys.argv) > 3 and sys.argv[3] == "sweep":
        for lanes in (8, 16, 32):
            for inl in (0, 4, 8):
                for minb in (3, 4):
                    os.environ.update(NP2_INFLATE_LANES=str(lanes), NP2_INFLATE_INLINE=str(inl), NP2_INFLATE_MINB=str(minb))
                    ks = [np2.bgzf_inflate(ctx, pin_in, po, pl, iz, out=pin_out)[1] for _ in range(4)]
                    assert bytes(pin_out.array[:total]) == want
                    print("  level %d sweep: lanes %2d inline<=%d min CTAs/SM %d: kernel %.2f ms" % (level, lanes, inl, minb, min(ks[1:])),
                          flush=True)
        for k in ("NP2_INFLATE_LANES", "NP2_INFLATE_INLINE", "NP2_INFLATE_MINB"):
            os.environ.pop(k, None)
    print("level %d: %d members, %.1f MB -> %.1f MB | zlib on %d host threads %.1f ms (%.2f GB/s out) | device kernel %.2f ms "
          "(%.1f GB/s out, %.1f GB/s in) | call incl. H2D of the compressed span + D2H of the records: pinned source %.1f ms, "
          "pageable source %.1f ms | identical to zlib" % (
              level, len(po), len(buf) / 1e6, total / 1e6, cores, t_host * 1e3, total / 1e9 / t_host, res["pinned"][0],
              total / 1e6 / res["pinned"][0], len(buf) / 1e6 / res["pinned"][0], res["pinned"][1], res["pageable"][1]), flush=True)
