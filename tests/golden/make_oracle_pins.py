"""Writes tests/golden/oracle_pins.json: sha256 of the oracle's FASTA on seeded datasets (self-pin)."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import common  # noqa: E402
import oracle as O  # noqa: E402

pins = {}
for name in ["tiny20k", "clip120k"]:
    ds = common.dataset(name)
    j = O.Job(ds["contig"], ds["bam"], common.oracle_tables(ds), O.Opts(min_ctg_len=0))
    pos, base = j.consensus()
    pins[name] = hashlib.sha256(O.format_fasta(name, pos, base)).hexdigest()
json.dump(pins, open(os.path.join(HERE, "oracle_pins.json"), "w"), indent=1)
print(pins)
