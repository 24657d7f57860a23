"""Developer aid: K5 probe rate vs the random-gather rate (optionally under NP2_L2_FETCH_GRANULARITY)."""
import sys
sys.path.insert(0, ".")
import json, torch
import bench
import nextpolish2_b200 as np2
ctx = np2.Context(0)
print(json.dumps(bench.yak_bench(ctx, np2, torch, 6550.1)))
