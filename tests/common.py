"""Shared helpers: seeded datasets and oracle/product comparison (tests only)."""
import functools

import numpy as np

import oracle as O  # oracle/oracle.py — test infrastructure
from nextpolish2_b200 import synth

KS = (21, 31)


@functools.lru_cache(maxsize=None)
def dataset(name):
    """name -> dict(contig, bam, haps, tables {k: (hashes, counts)})."""
    spec = {
        # haploid, assembly errors to fix
        "hap300k": dict(seed=101, L=300_000, depth=30, asm_err=2e-4, het=0.0),
        # diploid >= 500 kb: phasing, clipped reads get labelled, low-MAPQ and supplementary records filtered
        "dip600k": dict(seed=202, L=600_000, depth=40, asm_err=1e-4, het=0.002, frac_clip=0.03, frac_lowq=0.02, frac_supp=0.02),
        # < 500 kb with clipped reads (dropped), =/X cigars, noisier reads
        "clip120k": dict(seed=303, L=120_000, depth=25, asm_err=3e-4, het=0.001, frac_clip=0.05, eqx=True, read_err=0.004),
        # tandem repeats: long multi-entry runs in the DP
        "tandem200k": dict(seed=404, L=200_000, depth=30, asm_err=2e-4, het=0.001, tandem=0.08),
        # deep: the 60-candidate cap
        "deep80k": dict(seed=505, L=80_000, depth=90, asm_err=5e-4, het=0.002, mean_len=9000, sd_len=1500, min_len=3000),
        "tiny20k": dict(seed=606, L=20_000, depth=20, asm_err=5e-4, het=0.0, mean_len=6000, sd_len=1000, min_len=2000),
    }[name]
    s = dict(spec)
    seed, L = s.pop("seed"), s.pop("L")
    tandem = s.pop("tandem", 0.0)
    A = synth.genome(seed, L, tandem_frac=tandem)
    c = synth.make_contig(seed + 1, A, threads=4, **s)
    haps = [c["hap1"]] + ([c["hap2"]] if len(c["hap2"]) else [])
    tables = {k: synth.make_table(seed + 2, k, haps) for k in (21, 31, 51)}
    return {"contig": A, "bam": c["bam"], "haps": haps, "tables": tables, "n_reads": c["n_reads"]}


def oracle_tables(ds, ks=KS):
    return [O.Table.from_arrays(k, *ds["tables"][k]) for k in ks]


def gpu_tables(ctx, ds, ks=KS):
    import nextpolish2_b200 as np2
    return [np2.Table.from_arrays(ctx, k, *ds["tables"][k]) for k in ks]


def same_opts(**kw):
    import nextpolish2_b200 as np2
    kw.setdefault("min_ctg_len", 0)
    return O.Opts(**kw), np2.Opts(**kw)


def assert_same(name, a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    if a.shape != b.shape:
        n = min(len(a), len(b))
        first = int(np.argmax(a[:n] != b[:n])) if n and (a[:n] != b[:n]).any() else n
        raise AssertionError("%s: length %d (oracle) vs %d (gpu); first difference at %d" % (name, len(a), len(b), first))
    if not np.array_equal(a, b):
        idx = np.flatnonzero(a != b)
        i = int(idx[0])
        raise AssertionError("%s: %d of %d differ; first at %d: oracle %r gpu %r (context oracle %r gpu %r)" % (
            name, len(idx), len(a), i, a[i], b[i], a[max(0, i - 3):i + 4].tolist(), b[max(0, i - 3):i + 4].tolist()))


def assert_same_dict(prefix, da, db, keys=None):
    for k in keys or da.keys():
        assert_same(prefix + "." + k, da[k], db[k])
