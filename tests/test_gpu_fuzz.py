"""-m gpu: randomised sweep of the synthetic generator's parameters (depth, error rates, heterozygosity, read lengths,
clips / low MAPQ / supplementary fractions, tandem repeats, =/X CIGARs, options) - final consensus and, for the first
iteration, regions / candidates / dropped reads of the CUDA path against the oracle.  Seeds are fixed: a failure
reproduces."""
import numpy as np
import pytest

import common
import oracle as O
from nextpolish2_b200 import synth

pytestmark = pytest.mark.gpu


def draw(seed):
    rng = np.random.default_rng(9000 + seed)
    L = int(rng.integers(30_000, 160_000))
    mean_len = float(rng.integers(2500, 12000))
    p = dict(depth=float(rng.integers(6, 70)), asm_err=float(rng.choice([0, 1e-4, 5e-4, 2e-3])),
             het=float(rng.choice([0, 0, 5e-4, 3e-3, 1e-2])), read_err=float(rng.choice([5e-4, 2e-3, 8e-3])),
             mean_len=mean_len, sd_len=mean_len / 6, min_len=1500.0, max_len=3 * mean_len,
             frac_clip=float(rng.choice([0, 0.05])), frac_lowq=float(rng.choice([0, 0.05])),
             frac_supp=float(rng.choice([0, 0.05])), eqx=bool(rng.integers(2)))
    tandem = float(rng.choice([0, 0, 0.1]))
    opts = {}
    if rng.integers(3) == 0:
        opts["iter_count"] = int(rng.integers(1, 4))
    if rng.integers(4) == 0:
        opts["use_all_reads"] = 1
    if rng.integers(4) == 0:
        opts["model"] = 1
    if rng.integers(4) == 0:
        opts["min_kmer_count"] = int(rng.choice([1, 12, 30]))
    if rng.integers(4) == 0:
        opts["max_clip_len"] = 1000
    if rng.integers(5) == 0:
        opts["use_supplementary"] = 1
    ks = [(21, 31), (21,), (31,), (21, 31, 51), (17, 25)][int(rng.integers(5))]
    return L, p, tandem, opts, ks


@pytest.mark.parametrize("seed", range(48))
def test_random_configuration(ctx, seed):
    import nextpolish2_b200 as np2
    L, p, tandem, optkw, ks = draw(seed)
    A = synth.genome(7000 + seed, L, tandem_frac=tandem)
    c = synth.make_contig(8000 + seed, A, threads=4, **p)
    haps = [c["hap1"]] + ([c["hap2"]] if len(c["hap2"]) else [])
    tabs = {k: synth.make_table(8500 + seed, k, haps) for k in ks}
    oo, go = common.same_opts(**optkw)
    ot = [O.Table.from_arrays(k, *tabs[k]) for k in ks]
    gt = [np2.Table.from_arrays(ctx, k, *tabs[k]) for k in ks]
    oj = O.Job(A, c["bam"], ot, oo, dump_iter=0)
    gj = np2.Job(ctx, A, c["bam"], gt, go).upload().run(0)
    common.assert_same_dict("regions", oj.regions(), gj.regions(), ["start", "end", "lable"])
    common.assert_same_dict("cand", oj.candidates(), gj.candidates(), ["roff", "order", "seq_off", "seq", "kmer", "kscore"])
    common.assert_same("dropped", oj.dropped(), gj.dropped())
    opos, obase = oj.consensus()
    gpos, gbase = gj.consensus()
    common.assert_same("final.base", obase, gbase)
    common.assert_same("final.pos", opos, gpos)
    gj.destroy()
    # and once more without any dump (the sparse host view of the final phase)
    gpos2, gbase2 = np2.polish_contig(ctx, A, c["bam"], gt, go)
    common.assert_same("final.base (no dump)", obase, gbase2)
    common.assert_same("final.pos (no dump)", opos, gpos2)
