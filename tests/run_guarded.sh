#!/bin/bash
# Developer aid: run a command with a wall-clock limit and an RSS watchdog (kills only the PID it started),
# so a runaway host allocation cannot take the GPU box down.   usage: run_guarded.sh <seconds> <max_rss_gb> cmd...
limit=$1; shift
maxgb=$1; shift
"$@" &
pid=$!
( for ((i=0;i<limit;i++)); do
    sleep 1
    kill -0 $pid 2>/dev/null || exit 0
    rss=$(awk '/VmRSS/{print $2}' /proc/$pid/status 2>/dev/null || echo 0)
    if [ "${rss:-0}" -gt $((maxgb*1000000)) ]; then echo "[guard] RSS ${rss} kB > ${maxgb} GB: killing $pid"; kill -9 $pid; exit 0; fi
  done
  echo "[guard] time limit ${limit}s: killing $pid"; kill -9 $pid ) &
wd=$!
wait $pid
rc=$?
kill $wd 2>/dev/null
exit $rc
