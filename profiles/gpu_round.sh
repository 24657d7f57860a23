#!/bin/bash
# One GPU-box visit: parity tests, bench lines (haploid configs[1], one diploid contig of configs[2], reference arm),
# ncu launch list and a full capture of the pipeline's kernels at the bench's own size.
# usage (under gpurun): bash profiles/gpu_round.sh <tag>
tag=${1:-run}
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest.log 2>&1
tail -3 gpurun_out/${tag}_pytest.log
python bench.py --steps 50 --warmup 6 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_bench.json; tail -2 gpurun_out/${tag}_bench.err
python bench.py --het 0.01 --steps 15 --warmup 4 --no-yak-bench --cpu-steps 3 > gpurun_out/${tag}_bench_diploid.json 2>> gpurun_out/${tag}_bench.err
cut -c1-400 gpurun_out/${tag}_bench_diploid.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-yak-bench --e2e-inflight 1 > gpurun_out/${tag}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:'k_pack_columns|k_trim_scan|k_pileup_emit|k_emit_singles|k_emit_runs|k_cand_write|k_pair_scan|k_dp_runs|k_region_seed|k_region_hete|k_region_select|k_gather_seq_tma|k_pos_finalize' \
    -c 14 -o gpurun_out/${tag}_full python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-yak-bench --e2e-inflight 1 \
    > gpurun_out/${tag}_ncu_full.log 2>&1
ls -la gpurun_out | tail -10
