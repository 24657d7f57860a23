"""CPU: the C-ABI library loads, exports every symbol include/np2gpu.h declares, and fails loudly without a GPU.
No compute entry point is called here."""
import os
import re

import numpy as np
import pytest

import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "np2gpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(np2_[a-z0-9_]+)\s*\(", txt)))


def test_exports_match_header():
    import nextpolish2_b200 as np2
    L = np2.load_library()
    decl = _declared()
    assert len(decl) >= 25
    missing = [n for n in decl if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(np2.api.EXPORTS) == decl  # the ctypes mirror binds exactly the declared ABI


def test_opts_layout_and_defaults():
    import ctypes as C
    import nextpolish2_b200 as np2
    o = np2.Opts(min_kmer_count=0, iter_count=9)
    np2.load_library().np2_opts_default(C.byref(o))
    d = np2.Opts()
    for f, _ in np2.Opts._fields_:
        assert getattr(o, f) == getattr(d, f), f
    assert C.sizeof(np2.Opts) == C.sizeof(O.Opts) == 72
    assert (d.min_kmer_count, d.iter_count, d.min_read_len, d.min_ctg_len, d.max_indel_len) == (5, 2, 1000, 1000000, 20)
    assert (d.min_map_len, d.min_map_fra, d.min_map_qual, d.max_clip_len) == (500, 0.5, 1, 100)


def test_format_fasta_matches_oracle():
    import nextpolish2_b200 as np2
    pos = np.array([0, 1, 1, 1, 7], np.uint32)
    base = np.frombuffer(b"AcgTn", np.uint8)
    for up in (False, True):
        for op in (False, True):
            assert np2.format_fasta("ctg 1", pos, base, up, op) == O.format_fasta("ctg 1", pos, base, up, op)


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point must fail loudly (no silent CPU path)."""
    import torch
    import nextpolish2_b200 as np2
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(np2.Np2Error) as e:
        np2.Context(0)
    assert e.value.code == -1 and "no CPU fallback" in str(e.value)


def test_product_does_not_touch_oracle():
    """The oracle is test infrastructure: nothing under nextpolish2_b200/ may import, link or open it."""
    pkg = os.path.join(ROOT, "nextpolish2_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                for needle in ("np2_oracle", "libnp2oracle", "import oracle", "from oracle", "np2o_"):
                    assert needle not in src, (f, needle)
    import subprocess
    out = subprocess.run(["ldd", os.path.join(pkg, "libnp2gpu.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out
