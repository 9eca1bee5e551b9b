#!/bin/bash
# Round 2, call C: A/B matrix of compile-time variants (thread order, fast pair math, L1 tensor prefetch) + failing tests.
# MEASUREMENT infrastructure, not part of the product.
set -u
OUT=gpurun_out/${1:-r2c}
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/gpu.txt" 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cold_calls_and_decoupled_gravity.py -m gpu -q -p no:cacheprovider > "$OUT/pytest_gpu.log" 2>&1
echo "pytest rc=$?"; tail -n 8 "$OUT/pytest_gpu.log"
run_variant() {  # <name> <workload> <flags>
    B200SPH_EXTRA_FLAGS="$3" python -m miluphcuda_b200.build $2 --force > "$OUT/build_$1_$2.log" 2>&1 || { echo "build failed $1 $2"; tail -3 "$OUT/build_$1_$2.log"; }
    grep -h "k_forces\|registers" "$OUT/build_$1_$2.log" | head -0
    timeout 300 python bench.py --workload $2 --state step0 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > "$OUT/bench_$1_$2.json" 2> "$OUT/bench_$1_$2.err"
    echo "== $1 $2 ('$3') rc=$?"
    python tools/show_bench.py "$OUT/bench_$1_$2.json"
}
for w in impact sedov nakamura giant_hydro; do
    run_variant base $w ""
    run_variant fast $w "-DB200_FAST_PAIR_MATH=1"
    run_variant brickfast $w "-DB200_FAST_PAIR_MATH=1 -DB200_BRICK_ORDER=1"
done
run_variant fastpf impact "-DB200_FAST_PAIR_MATH=1 -DB200_PREFETCH_TENSORS_L1=1"
run_variant fastpf nakamura "-DB200_FAST_PAIR_MATH=1 -DB200_PREFETCH_TENSORS_L1=1"
B200SPH_EXTRA_FLAGS="-DB200_FAST_PAIR_MATH=1" python -m miluphcuda_b200.build --force > /dev/null 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider > "$OUT/pytest_fast_all.log" 2>&1
echo "fast-math parity (all configs) rc=$?"; tail -n 5 "$OUT/pytest_fast_all.log"
