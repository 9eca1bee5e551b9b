#!/bin/bash
# Round 2, call B: GPU tests on the brick-ordered kernels, step-0 benches, fast-pair-math variant, ncu captures.
# MEASUREMENT infrastructure, not part of the product.
set -u
OUT=gpurun_out/${1:-r2b}
PH=${2:-tests,bench,variants,ncu}
has() { [[ ",$PH," == *",$1,"* ]]; }
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/gpu.txt" 2>&1
if has tests; then
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x --durations=8 > "$OUT/pytest_gpu.log" 2>&1
echo "pytest rc=$?"; tail -n 25 "$OUT/pytest_gpu.log"
fi
if has bench; then
for w in impact sedov rings giant_hydro nakamura; do
    timeout 300 python bench.py --workload $w --state step0 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > "$OUT/bench_$w.json" 2> "$OUT/bench_$w.err"
    echo "bench $w rc=$?"; tail -n 2 "$OUT/bench_$w.err"
done
python tools/show_bench.py "$OUT"/bench_impact.json "$OUT"/bench_sedov.json "$OUT"/bench_rings.json "$OUT"/bench_giant_hydro.json "$OUT"/bench_nakamura.json
fi
if has variants; then
for w in impact sedov; do
    B200SPH_EXTRA_FLAGS="-DB200_FAST_PAIR_MATH=1" python -m miluphcuda_b200.build $w --force > "$OUT/build_fast_$w.log" 2>&1
    timeout 300 python bench.py --workload $w --state step0 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > "$OUT/bench_fast_$w.json" 2> "$OUT/bench_fast_$w.err"
    echo "fast math $w rc=$?"
    python tools/show_bench.py "$OUT/bench_fast_$w.json"
    timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "$w" > "$OUT/pytest_fast_$w.log" 2>&1
    echo "fast math parity $w rc=$?"; tail -n 3 "$OUT/pytest_fast_$w.log"
    python -m miluphcuda_b200.build $w --force > /dev/null 2>&1
done
fi
if has ncu; then
for w in sedov impact; do
    timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file "$OUT/launches_$w.csv" \
        python bench.py --workload $w --state step0 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > "$OUT/ncu_launch_$w.log" 2>&1
    echo "ncu launches $w rc=$?"
    timeout 700 ncu --set full --clock-control none --import-source on -k 'regex:k_forces|k_neighbours|k_density|k_correction|k_pointwise' -s 12 -c 5 -f -o "$OUT/full_$w" \
        python bench.py --workload $w --state step0 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > "$OUT/ncu_full_$w.log" 2>&1
    echo "ncu full $w rc=$?"
done
fi
ls "$OUT"
