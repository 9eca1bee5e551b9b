#!/bin/bash
# One short gpurun call: GPU parity tests, then our bench arm on the named workloads.  MEASUREMENT infrastructure.
#   gpurun --timeout 1200 -- 'bash tools/gpu_quick.sh <tag> "sedov impact" [pytest-args]'
set -u
TAG=${1:-quick}
WORKLOADS=${2:-sedov impact}
PYTEST_ARGS=${3:-}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/gpu.txt" 2>&1
timeout 900 python -m pytest tests -m gpu -x -q $PYTEST_ARGS > "$OUT/pytest_gpu.log" 2>&1
echo "pytest rc=$?" | tee -a "$OUT/pytest_gpu.log"
tail -n 15 "$OUT/pytest_gpu.log"
for w in $WORKLOADS; do
    extra="--no-cpu-baseline"
    [ "$w" = sedov ] && extra=""
    timeout 300 python bench.py --workload $w --steps 10 --warmup 3 $extra > "$OUT/bench_$w.json" 2> "$OUT/bench_$w.err"
    echo "bench $w rc=$?"
    tail -n 3 "$OUT/bench_$w.err"
done
python tools/show_bench.py $(for w in $WORKLOADS; do echo "$OUT/bench_$w.json"; done)
