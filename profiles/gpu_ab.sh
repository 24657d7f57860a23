#!/bin/bash
# One short GPU-box visit: kernel-variant A/B on the bench workload, then the parity tests with the built-in defaults,
# then the fuzz parity tests under the other batched variant.   usage (under gpurun): bash profiles/gpu_ab.sh <tag>
tag=${1:-ab}
mkdir -p gpurun_out
timeout 150 python profiles/ab_kernels.py 15 > gpurun_out/${tag}_ab.log 2>&1
cat gpurun_out/${tag}_ab.log | tail -9
( time timeout 150 python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest.log 2>&1
tail -5 gpurun_out/${tag}_pytest.log
( NP2_PILE_BATCH=2 timeout 60 python -m pytest tests/test_gpu_fuzz.py -m gpu -x -q ) > gpurun_out/${tag}_pytest_pile2.log 2>&1
tail -2 gpurun_out/${tag}_pytest_pile2.log
timeout 120 python bench.py --steps 30 --warmup 6 --no-yak-bench --cpu-steps 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_bench.json | cut -c1-1200
