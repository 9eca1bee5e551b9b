#!/bin/bash
# ncu captures for one workload: launch list + full capture of named kernels.  MEASUREMENT infrastructure.
#   gpurun --timeout 1200 -- 'bash tools/gpu_ncu.sh <tag> <workload> <kernel-regex> [skip] [count]'
set -u
TAG=${1:-ncu}; W=${2:-sedov}; K=${3:-k_forces}; SKIP=${4:-0}; COUNT=${5:-2}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file "$OUT/launches_$W.csv" \
    python bench.py --workload $W --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > "$OUT/ncu_launch_$W.log" 2>&1
echo "ncu launches $W rc=$?"
timeout 700 ncu --set full --clock-control none --import-source on -k "regex:$K" -s $SKIP -c $COUNT -f -o "$OUT/full_$W" \
    python bench.py --workload $W --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > "$OUT/ncu_full_$W.log" 2>&1
echo "ncu full $W rc=$?"
ls -la "$OUT"
