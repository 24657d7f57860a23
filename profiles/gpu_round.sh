#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list and a full capture of the pipeline's top kernels.
# usage (under gpurun): bash profiles/gpu_round.sh <tag>
tag=${1:-run}
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest.log 2>&1
tail -3 gpurun_out/${tag}_pytest.log
python bench.py --steps 30 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_bench.json
python bench.py --impl reference --steps 3 --warmup 0 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-yak-bench > gpurun_out/${tag}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:'k_expand|k_pack_columns|k_trim|k_cand_write|k_emit_singles|k_pair_scan|k_pileup_emit|k_pileup_count|k_dp_runs|k_region_seed|k_region_hete' \
    -c 14 -o gpurun_out/${tag}_full python bench.py --steps 1 --warmup 0 --length 4000000 --no-cpu-baseline --no-yak-bench \
    > gpurun_out/${tag}_ncu_full.log 2>&1
ls -la gpurun_out
