#!/bin/bash
# Round 2, call E: full GPU test suite after the refactors (staged evaluation, 8-way gravity walk, integrator), gravity benches.
set -u
OUT=gpurun_out/${1:-r2e}
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/gpu.txt" 2>&1
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 > "$OUT/pytest_gpu.log" 2>&1
echo "pytest rc=$?"; grep -n "^E  \|passed\|failed" "$OUT/pytest_gpu.log" | head -40
for w in giant_hydro giant_solid; do
    timeout 300 python bench.py --workload $w --state step0 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > "$OUT/bench_$w.json" 2> "$OUT/bench_$w.err"
    echo "bench $w rc=$?"; tail -n 2 "$OUT/bench_$w.err"
done
python tools/show_bench.py "$OUT"/bench_giant_hydro.json "$OUT"/bench_giant_solid.json
timeout 300 python bench.py --impl reference --workload giant_hydro --state step0 --steps 5 --warmup 2 > "$OUT/bench_ref_giant_hydro.json" 2> "$OUT/bench_ref_giant_hydro.err"
cut -c1-400 "$OUT/bench_ref_giant_hydro.json"
