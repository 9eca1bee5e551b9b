#!/bin/bash
# Round 2, call D: device integrator + reorder tests, benches with and without the persistent cell order.
# MEASUREMENT infrastructure, not part of the product.
set -u
OUT=gpurun_out/${1:-r2d}
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/gpu.txt" 2>&1
timeout 1200 python -m pytest tests/test_gpu_reorder.py tests/test_gpu_rk2_integrator.py -m gpu -q -p no:cacheprovider > "$OUT/pytest_new.log" 2>&1
echo "pytest new rc=$?"; grep -n "^E  \|passed\|failed" "$OUT/pytest_new.log" | head -40
for w in impact sedov nakamura; do
    timeout 300 python bench.py --workload $w --state step0 --steps 10 --warmup 3 --no-cpu-baseline > "$OUT/bench_$w.json" 2> "$OUT/bench_$w.err"
    echo "bench $w rc=$?"; tail -n 2 "$OUT/bench_$w.err"
    timeout 300 python bench.py --workload $w --state step0 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-reorder > "$OUT/bench_noreorder_$w.json" 2> "$OUT/bench_noreorder_$w.err"
done
python tools/show_bench.py "$OUT"/bench_impact.json "$OUT"/bench_noreorder_impact.json "$OUT"/bench_sedov.json "$OUT"/bench_noreorder_sedov.json "$OUT"/bench_nakamura.json "$OUT"/bench_noreorder_nakamura.json
