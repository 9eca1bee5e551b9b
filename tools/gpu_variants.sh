#!/bin/bash
# Compare compile-time variants of one workload on the GPU box (nvcc is in the image).  MEASUREMENT infrastructure.
#   gpurun -- 'bash tools/gpu_variants.sh <tag> <workload> "<flags A>" "<flags B>" ...'
set -u
TAG=$1; W=$2; shift 2
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
i=0
for flags in "$@"; do
    B200SPH_EXTRA_FLAGS="$flags" python -m miluphcuda_b200.build $W --force > "$OUT/build_$i.log" 2>&1 || { echo "build failed: $flags"; tail -5 "$OUT/build_$i.log"; }
    timeout 300 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > "$OUT/bench_${W}_$i.json" 2> "$OUT/bench_${W}_$i.err"
    echo "== variant $i: '$flags' rc=$?"
    python tools/show_bench.py "$OUT/bench_${W}_$i.json"
    i=$((i+1))
done
