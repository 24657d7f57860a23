"""Host half of the phasing step (main.rs:994-1015 + louvain.rs:59-356) through the np2_debug_phase test seam against the
oracle's maps-of-maps restatement, on synthetic read x read agreement graphs: two haplotypes, overlapping reads,
het sites shared by overlapping reads, noise, reads that disagree with the ref read."""
import numpy as np
import pytest

import oracle as O
import nextpolish2_b200.api as api


def make_graph(seed, n_reads, span=40, noise=0.02, gap_every=0, ref=True):
    """keys (a << 32 | b, sorted), vals (#agree + #differ * (2^32 - 1)) of a two-haplotype read set"""
    rng = np.random.default_rng(seed)
    hap = rng.integers(0, 2, n_reads + 1)
    keys, vals = [], []
    if ref:
        for b in range(1, n_reads + 1):
            if rng.random() < 0.9:
                sites = int(rng.integers(1, 60))
                bad = int(rng.binomial(sites, noise if hap[b] == 0 else 1 - noise))
                keys.append(b)
                vals.append((sites - bad) + bad * ((1 << 32) - 1))
    for a in range(1, n_reads + 1):
        for b in range(a + 1, min(n_reads, a + span) + 1):
            if gap_every and (a // gap_every) != (b // gap_every):
                continue  # phase-block boundary: no shared het site
            sites = int(rng.integers(0, max(1, (span - (b - a)) * 2)))
            if sites == 0 and rng.random() < 0.7:
                continue
            bad = int(rng.binomial(sites, noise if hap[a] == hap[b] else 1 - noise))
            keys.append((a << 32) | b)
            vals.append((sites - bad) + bad * ((1 << 32) - 1))
    keys = np.array(keys, np.uint64)
    vals = np.array(vals, np.int64)
    o = np.argsort(keys, kind="stable")
    return keys[o], vals[o]


CASES = [dict(seed=1, n_reads=60), dict(seed=2, n_reads=400, span=30), dict(seed=3, n_reads=1500, span=50, noise=0.05),
         dict(seed=4, n_reads=1200, span=35, gap_every=150), dict(seed=5, n_reads=800, span=25, noise=0.2),
         dict(seed=6, n_reads=700, span=30, ref=False), dict(seed=7, n_reads=900, span=12, noise=0.35, gap_every=90)]


@pytest.mark.parametrize("case", CASES, ids=[str(c["seed"]) for c in CASES])
@pytest.mark.parametrize("model,use_all", [(0, False), (1, False), (0, True)])
def test_phase_matches_oracle(case, model, use_all):
    keys, vals = make_graph(**case)
    want = O.debug_phase(keys, vals, model, use_all)
    got = api.debug_phase(keys, vals, model, use_all)
    assert np.array_equal(np.sort(got), want)


def test_phase_empty_and_tiny():
    e = np.empty(0, np.uint64)
    assert len(api.debug_phase(e, e.astype(np.int64))) == 0 and len(O.debug_phase(e, e.astype(np.int64))) == 0
    keys = np.array([(1 << 32) | 2, (1 << 32) | 3, (2 << 32) | 3], np.uint64)
    vals = np.array([5, 3 * ((1 << 32) - 1), 4 * ((1 << 32) - 1) + 1], np.int64)
    assert np.array_equal(np.sort(api.debug_phase(keys, vals)), O.debug_phase(keys, vals))


def gadget(base=0):
    """A community that ends up with a negative internal weight and falls apart again (louvain.rs:136-165): 1 and 2
    join 3's community, then 3 leaves for 4; what is left of community 3 is {1, 2} with weight -5."""
    D = (1 << 32) - 1
    e = {(1, 2): 5 * D, (1, 3): 10, (1, 4): 100 * D, (2, 3): 8, (2, 4): 100 * D, (3, 4): 50}
    keys = np.array([((a + base) << 32) | (b + base) for a, b in e], np.uint64)
    return keys, np.array(list(e.values()), np.int64)


def test_phase_decluster_goes_through_general_path():
    keys, vals = gadget()
    got, path = api.debug_phase(keys, vals, with_path=True)
    assert path == 2
    assert np.array_equal(np.sort(got), O.debug_phase(keys, vals))
    # the same gadget inside a larger graph, for every model
    k2, v2 = make_graph(seed=11, n_reads=500, span=20)
    gk, gv = gadget(base=600)
    keys = np.concatenate([k2, gk])
    vals = np.concatenate([v2, gv])
    o = np.argsort(keys, kind="stable")
    keys, vals = keys[o], vals[o]
    for model, use_all in ((0, False), (1, False), (0, True)):
        got, path = api.debug_phase(keys, vals, model, use_all, with_path=True)
        assert path == 2
        assert np.array_equal(np.sort(got), O.debug_phase(keys, vals, model, use_all))
    assert api.debug_phase(k2, v2, with_path=True)[1] == 1  # the common case stays on the flat-array path


_CHILD = r"""
import json, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/oracle"); sys.path.insert(0, %(root)r + "/tests")
import numpy as np
import oracle as O
import nextpolish2_b200.api as api
from test_phase_host import make_graph
out = []
for case in %(cases)r:
    keys, vals = make_graph(**case)
    for model, use_all in ((0, False), (1, False), (0, True)):
        got, path = api.debug_phase(keys, vals, model, use_all, with_path=True)
        want = O.debug_phase(keys, vals, model, use_all)
        out.append([bool(np.array_equal(np.sort(got), want)), int(path), np.sort(got).tolist()])
print(json.dumps(out))
"""

PAR_CASES = [dict(seed=21, n_reads=8000, span=30, gap_every=150), dict(seed=22, n_reads=3000, span=40, gap_every=400, noise=0.1),
             dict(seed=23, n_reads=2500, span=25), dict(seed=24, n_reads=1500, span=20, gap_every=60, noise=0.3, ref=False)]


@pytest.mark.timeout(600)
def test_phase_concurrent_paths_match_serial_and_oracle():
    """the concurrent vertex-move path (closed id intervals on the host pool) and the concurrent aggregation
    only start above 100k adjacency entries.  A child process lowers the threshold (NP2_PHASE_PAR_MIN=1), uses 8 host
    threads and switches the invariant checks on (NP2_PHASE_CHECK=1: neighbour lists ascending, no self loops); its
    results must equal the oracle's AND the serial results of this process, graph by graph."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    serial = []
    for case in PAR_CASES:
        keys, vals = make_graph(**case)
        for model, use_all in ((0, False), (1, False), (0, True)):
            serial.append(np.sort(api.debug_phase(keys, vals, model, use_all)).tolist())
    env = dict(os.environ, NP2_PHASE_PAR_MIN="1", NP2_HOST_THREADS="8", NP2_PHASE_CHECK="1")
    r = subprocess.run([sys.executable, "-c", _CHILD % {"root": root, "cases": PAR_CASES}], env=env, capture_output=True,
                       text=True, timeout=550)
    assert r.returncode == 0, r.stderr[-2000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert len(res) == len(serial)
    for (ok, path, got), want in zip(res, serial):
        assert ok, "concurrent phasing differs from the oracle"
        assert path == 1
        assert got == want, "concurrent phasing differs from the serial run"
