#!/bin/bash
# Round 2, call K: larger single-GPU sets (the per-GPU share of the 16M / 8-GPU impact line, the largest impact set
# generated here, a 4M self-gravitating set): step-0 state, no host-buffer leg.
set -u
OUT=gpurun_out/${1:-r2k}
mkdir -p "$OUT"
run() {  # <workload> <particles>
    timeout 500 python bench.py --workload $1 --particles $2 --state step0 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > "$OUT/bench_$1_$2.json" 2> "$OUT/bench_$1_$2.err"
    echo "== $1 $2 rc=$?"; python tools/show_bench.py "$OUT/bench_$1_$2.json"; nvidia-smi --query-gpu=memory.used --format=csv,noheader
}
run impact 2000000
run giant_hydro 4000000
run impact 8000000
