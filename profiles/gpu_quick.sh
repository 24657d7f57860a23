#!/bin/bash
# Short GPU-box visit: parity tests + bench line.   usage (under gpurun): bash profiles/gpu_quick.sh <tag> [bench args]
tag=${1:-run}; shift
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest.log 2>&1
tail -5 gpurun_out/${tag}_pytest.log
python bench.py "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
