#!/bin/bash
# Round 2, call H: the team force kernel (several lanes per particle) against one particle per lane, team sizes,
# register budgets of the team kernel, the leaner gravity walk; parity of all of it; ncu of g_walk and k_forces_team.
set -u
OUT=gpurun_out/${1:-r2h}
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > "$OUT/gpu.txt"
run() {  # <tag> <workload> <env...>
    local tag=$1 w=$2; shift 2
    env "$@" timeout 300 python bench.py --workload $w --state step0 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > "$OUT/bench_${tag}_$w.json" 2> "$OUT/bench_${tag}_$w.err"
    echo "== $tag $w rc=$?"
    python tools/show_bench.py "$OUT/bench_${tag}_$w.json"
}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_c_host_example.py -m gpu -q -x -p no:cacheprovider > "$OUT/pytest_parity.log" 2>&1
echo "parity rc=$?"; tail -n 4 "$OUT/pytest_parity.log"
for w in impact nakamura sedov; do
    run team4 $w B200SPH_PAIR_TEAMS=7
    run lane $w B200SPH_PAIR_TEAMS=0
    for v in team2 team8; do run $v $w B200SPH_LIBDIR=$PWD/miluphcuda_b200/lib_$v; done
done
for w in impact nakamura; do
    for v in tsmem tsmem4 tcap4; do run $v $w B200SPH_LIBDIR=$PWD/miluphcuda_b200/lib_$v; done
done
for cs in 0.5 0.7; do run cell$cs impact B200SPH_CELL_SCALE=$cs; done
run cell0.8 sedov B200SPH_CELL_SCALE=0.8
run cell1.25 sedov B200SPH_CELL_SCALE=1.25
run team4 giant_hydro B200SPH_PAIR_TEAMS=7
run team4 giant_solid B200SPH_PAIR_TEAMS=7
run lane giant_solid B200SPH_PAIR_TEAMS=0
# ncu: gravity walk and the team kernel (one launch each, warm)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:g_walk --launch-skip 3 --launch-count 1 -f -o "$OUT/ncu_g_walk" \
    python bench.py --workload giant_hydro --state step0 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > "$OUT/ncu_g_walk.log" 2>&1
echo "ncu g_walk rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_forces|k_correction|k_density" --launch-skip 6 --launch-count 2 -f -o "$OUT/ncu_forces_team_impact" \
    python bench.py --workload impact --state step0 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > "$OUT/ncu_forces_team.log" 2>&1
echo "ncu k_forces_team rc=$?"
B200SPH_PAIR_TEAMS=0 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_forces|k_correction|k_density" --launch-skip 6 --launch-count 2 -f -o "$OUT/ncu_forces_lane_impact" \
    python bench.py --workload impact --state step0 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > "$OUT/ncu_forces_lane.log" 2>&1
echo "ncu k_forces lane rc=$?"
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider --deselect tests/test_gpu_parity.py > "$OUT/pytest_gpu.log" 2>&1
echo "pytest gpu rc=$?"; tail -n 6 "$OUT/pytest_gpu.log"
ls -la "$OUT"/*.ncu-rep
