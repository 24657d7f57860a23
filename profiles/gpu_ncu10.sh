#!/bin/bash
# ncu --set full of the streaming kernels at the bench's own launch size (10 Mbp / 30x), plus the launch list.
tag=${1:-run}
mkdir -p gpurun_out
python bench.py --steps 30 --warmup 6 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_bench.json; tail -2 gpurun_out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-yak-bench --e2e-inflight 1 > gpurun_out/${tag}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:'k_pack_columns|k_trim_scan|k_pileup_emit|k_pileup_count|k_emit_singles|k_cand_write|k_pair_scan|k_dp_runs|k_region_seed|k_region_hete|k_region_select|k_table_probe' \
    -c 12 -o gpurun_out/${tag}_full python bench.py --steps 1 --warmup 0 --no-cpu-baseline --e2e-inflight 1 \
    > gpurun_out/${tag}_ncu_full.log 2>&1
ls -la gpurun_out | tail -8
