#!/bin/bash
# Profile visit (round 2 pipeline): launch list + ncu --set full of the pipeline kernels at the bench's own size,
# converted to CSV / per-line stall summaries ON THE BOX (the reports themselves are too large to travel back).
# usage (under gpurun): bash profiles/gpu_prof.sh <tag> [bench args for the captured run]
tag=${1:-prof}; shift
mkdir -p gpurun_out /tmp/np2prof
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-yak-bench --no-extras --no-strong --no-verify --e2e-inflight 1 "$@" \
    > gpurun_out/${tag}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'k_pileup_stripe|k_block_flags|k_dp_runs|k_pack_columns|k_scan_apply|k_region_hete|k_pair_scan|k_cand_write|k_region_select|k_region_seed|k_emit_runs|k_emit_singles|k_trim_scan|k_seed_gather|k_assemble|k_consensus' \
    -c 24 -o /tmp/np2prof/${tag}_full python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-yak-bench --no-extras --no-strong --no-verify --e2e-inflight 1 "$@" \
    > gpurun_out/${tag}_ncu_full.log 2>&1
ncu -i /tmp/np2prof/${tag}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
python profiles/source_hotspots.py /tmp/np2prof/${tag}_full.ncu-rep 10 > gpurun_out/${tag}_source_hotspots.txt 2>&1
ls -la /tmp/np2prof gpurun_out | tail -12
