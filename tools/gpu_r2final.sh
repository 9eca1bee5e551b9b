#!/bin/bash
# Round 2, final one-GPU measurement: parity tests, smoke, both bench arms on the default workload (impact, evolved
# state), the other workloads, the last register A/B of the solid force loop, ncu launch lists and full captures.
#   gpurun --timeout 2400 -- 'bash tools/gpu_r2final.sh <tag>'          MEASUREMENT infrastructure.
set -u
TAG=${1:-r2final}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/gpu.txt" 2>&1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=10 > "$OUT/pytest_gpu.log" 2>&1
echo "pytest rc=$?"; tail -n 4 "$OUT/pytest_gpu.log"
python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?"; tail -n 2 "$OUT/smoke.log"
# the driver's two commands, as the driver runs them
timeout 600 python bench.py --impl reference > "$OUT/bench_reference_impact.json" 2> "$OUT/bench_reference_impact.err"; echo "ref impact rc=$?"
timeout 600 python bench.py > "$OUT/bench_impact.json" 2> "$OUT/bench_impact.err"; echo "bench impact rc=$?"
for w in sedov giant_hydro; do
    timeout 600 python bench.py --impl reference --workload $w > "$OUT/bench_reference_$w.json" 2> "$OUT/bench_reference_$w.err"; echo "ref $w rc=$?"
done
for w in sedov nakamura giant_hydro giant_solid rings shocktube; do
    timeout 600 python bench.py --workload $w --no-cpu-baseline > "$OUT/bench_$w.json" 2> "$OUT/bench_$w.err"; echo "bench $w rc=$?"
done
python tools/show_bench.py "$OUT"/bench_impact.json "$OUT"/bench_sedov.json "$OUT"/bench_nakamura.json "$OUT"/bench_giant_hydro.json "$OUT"/bench_giant_solid.json "$OUT"/bench_rings.json "$OUT"/bench_shocktube.json
grep -h '"impl": "reference"' "$OUT"/bench_reference_*.json | python -c "
import sys, json
for line in sys.stdin:
    d = json.loads(line); print('reference', d['config']['workload'][:40], 'value=%.4g ms/step=%.4g' % (d['value'], d['ms_per_step']))"
# register budget of the 3-D solid force loop (own tensors in shared memory, 128-register caps): step-0 state, no e2e
for v in smem smem8 cap8; do
    B200SPH_LIBDIR=$PWD/miluphcuda_b200/lib_$v timeout 300 python bench.py --workload impact --state step0 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e \
        > "$OUT/bench_${v}_impact.json" 2> "$OUT/bench_${v}_impact.err"
    echo "== $v rc=$?"; python tools/show_bench.py "$OUT/bench_${v}_impact.json"
done
timeout 300 python bench.py --workload impact --state step0 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > "$OUT/bench_step0_impact.json" 2> "$OUT/bench_step0_impact.err"
python tools/show_bench.py "$OUT/bench_step0_impact.json"
# ncu (step-0 state: the evolved state is prepared by the reference binary, whose kernels would fill the launch list)
ncu_one() {  # workload regex skip count
    timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches_$1.csv" \
        python bench.py --workload $1 --state step0 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > "$OUT/ncu_launch_$1.log" 2>&1
    timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -f -o "$OUT/full_$1" \
        python bench.py --workload $1 --state step0 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > "$OUT/ncu_full_$1.log" 2>&1
    echo "ncu $1 rc=$?"
}
ncu_one impact "k_forces|k_neighbours|k_correction|k_pointwise" 12 4
ncu_one sedov "k_forces|k_neighbours|k_density" 9 3
ncu_one giant_hydro "g_walk|k_forces" 6 2
ls -la "$OUT" | head -60
