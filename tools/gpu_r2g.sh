#!/bin/bash
# Round 2, call G: A/B of the solid force kernel's register budget (own tensors in shared memory, launch bounds) and of
# running k_pointwise beside the search.  Variant libraries are prebuilt into miluphcuda_b200/lib_<name>/.
set -u
OUT=gpurun_out/${1:-r2g}
mkdir -p "$OUT"
run() {  # <tag> <workload> <env...>
    local tag=$1 w=$2; shift 2
    env "$@" timeout 300 python bench.py --workload $w --state step0 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > "$OUT/bench_${tag}_$w.json" 2> "$OUT/bench_${tag}_$w.err"
    echo "== $tag $w rc=$?"
    python tools/show_bench.py "$OUT/bench_${tag}_$w.json"
}
for w in impact nakamura; do
    run base $w B200SPH_OVERLAP_POINTWISE=1
    run nooverlap $w B200SPH_OVERLAP_POINTWISE=0
    for v in smem smem8 cap8; do
        run $v $w B200SPH_LIBDIR=$PWD/miluphcuda_b200/lib_$v
    done
done
run base giant_solid B200SPH_OVERLAP_POINTWISE=1
run smem8 giant_solid B200SPH_LIBDIR=$PWD/miluphcuda_b200/lib_smem8
B200SPH_LIBDIR=$PWD/miluphcuda_b200/lib_smem8 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "impact or nakamura or giant_solid" > "$OUT/pytest_smem8.log" 2>&1
echo "parity smem8 rc=$?"; tail -n 3 "$OUT/pytest_smem8.log"
