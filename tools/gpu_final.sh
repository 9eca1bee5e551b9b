#!/bin/bash
# Round-end measurement on ONE GPU: parity tests, both bench arms, ncu launch lists and full captures.
#   gpurun --timeout 1500 -- 'bash tools/gpu_final.sh <tag>'          MEASUREMENT infrastructure.
set -u
TAG=${1:-final}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/gpu.txt" 2>&1
timeout 900 python -m pytest tests -m gpu -q > "$OUT/pytest_gpu.log" 2>&1
echo "pytest rc=$?"; tail -n 3 "$OUT/pytest_gpu.log"
python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?"; tail -n 2 "$OUT/smoke.log"
timeout 300 python bench.py --steps 10 --warmup 3 > "$OUT/bench_sedov.json" 2> "$OUT/bench_sedov.err"; echo "bench sedov rc=$?"
for w in impact rings giant_hydro giant_solid shocktube; do
    timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > "$OUT/bench_$w.json" 2> "$OUT/bench_$w.err"; echo "bench $w rc=$?"
done
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > "$OUT/bench_ref_sedov.json" 2> "$OUT/bench_ref_sedov.err"; echo "ref sedov rc=$?"
timeout 300 python bench.py --impl reference --workload impact --steps 3 --warmup 2 > "$OUT/bench_ref_impact.json" 2> "$OUT/bench_ref_impact.err"; echo "ref impact rc=$?"
python tools/show_bench.py "$OUT"/bench_sedov.json "$OUT"/bench_impact.json "$OUT"/bench_rings.json "$OUT"/bench_giant_hydro.json "$OUT"/bench_giant_solid.json "$OUT"/bench_shocktube.json
ncu_one() {  # workload regex skip count
    timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file "$OUT/launches_$1.csv" \
        python bench.py --workload $1 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > "$OUT/ncu_launch_$1.log" 2>&1
    timeout 500 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -f -o "$OUT/full_$1" \
        python bench.py --workload $1 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > "$OUT/ncu_full_$1.log" 2>&1
    echo "ncu $1 rc=$?"
}
ncu_one sedov "k_forces|k_neighbours|k_density" 9 3
ncu_one impact "k_forces|k_neighbours|k_correction|k_pointwise" 12 4
ncu_one giant_hydro "g_walk|k_forces" 6 2
ls -la "$OUT" | head -50
