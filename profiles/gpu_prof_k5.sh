#!/bin/bash
# ncu --set full of K5 (k_table_probe) on the bench's 2^26-key table: DRAM bytes per probe.  usage: bash profiles/gpu_prof_k5.sh <tag>
tag=${1:-k5}
mkdir -p gpurun_out /tmp/np2prof
timeout 600 ncu --set full --clock-control none -k regex:'k_table_probe' -c 3 -o /tmp/np2prof/${tag}_probe \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --no-strong --no-verify --e2e-inflight 1 > gpurun_out/${tag}_ncu_probe.log 2>&1
ncu -i /tmp/np2prof/${tag}_probe.ncu-rep --page raw --csv > gpurun_out/${tag}_probe_raw.csv 2>/dev/null
ls -la /tmp/np2prof | tail -3
