#!/bin/bash
# Round 2, call A: new golden vectors from the reference, all GPU tests (no -x), impact/sedov bench both arms.
# MEASUREMENT infrastructure, not part of the product.
set -u
OUT=gpurun_out/r2a
mkdir -p "$OUT/golden"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/gpu.txt" 2>&1
timeout 600 python oracle/make_golden.py --out "$OUT/golden" \
    --configs nakamura,sedov_ignore,impact_ignore,giant_ignore,impact_crush1,impact_crush2,impact_crush3,impact_crush4 > "$OUT/golden.log" 2>&1
echo "golden rc=$?"; tail -n 12 "$OUT/golden.log"
cp "$OUT"/golden/*.npz tests/golden/ 2>/dev/null
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=15 > "$OUT/pytest_gpu.log" 2>&1
echo "pytest rc=$?"; tail -n 60 "$OUT/pytest_gpu.log"
timeout 600 python bench.py --steps 10 --warmup 3 > "$OUT/bench_impact.json" 2> "$OUT/bench_impact.err"
echo "bench impact rc=$?"; tail -n 3 "$OUT/bench_impact.err"
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > "$OUT/bench_ref_impact.json" 2> "$OUT/bench_ref_impact.err"
echo "bench ref impact rc=$?"; tail -n 3 "$OUT/bench_ref_impact.err"
timeout 300 python bench.py --workload sedov --steps 10 --warmup 3 --no-cpu-baseline > "$OUT/bench_sedov.json" 2> "$OUT/bench_sedov.err"
echo "bench sedov rc=$?"; tail -n 3 "$OUT/bench_sedov.err"
python tools/show_bench.py "$OUT"/bench_impact.json "$OUT"/bench_sedov.json
cut -c1-1500 "$OUT/bench_ref_impact.json"
