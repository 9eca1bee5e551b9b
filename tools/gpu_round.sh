#!/bin/bash
# One gpurun call: parity tests, both bench arms, launch list and ncu captures.  Everything lands in gpurun_out/<tag>/.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh <tag> [phases]'      phases default: tests,bench,ncu,drift
# MEASUREMENT infrastructure, not part of the product.
set -u
TAG=${1:-r01}
PHASES=${2:-tests,bench,ncu}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
has() { [[ ",$PHASES," == *",$1,"* ]]; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/gpu.txt" 2>&1

if has tests; then
    timeout 600 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1
    echo "pytest rc=$?" | tee -a "$OUT/pytest_gpu.log"
    tail -n 5 "$OUT/pytest_gpu.log"
fi

if has bench; then
    timeout 300 python bench.py --steps 10 --warmup 3 > "$OUT/bench_sedov.json" 2> "$OUT/bench_sedov.err"
    echo "bench sedov rc=$?"
    timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > "$OUT/bench_ref_sedov.json" 2> "$OUT/bench_ref_sedov.err"
    for w in impact rings giant_hydro; do
        timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > "$OUT/bench_$w.json" 2> "$OUT/bench_$w.err"
        echo "bench $w rc=$?"
    done
    timeout 300 python bench.py --impl reference --workload impact --steps 3 --warmup 2 > "$OUT/bench_ref_impact.json" 2> "$OUT/bench_ref_impact.err"
    python tools/show_bench.py "$OUT"/bench_sedov.json "$OUT"/bench_impact.json "$OUT"/bench_rings.json "$OUT"/bench_giant_hydro.json
fi

if has ncu; then
    for w in sedov impact; do
        timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file "$OUT/launches_$w.csv" \
            python bench.py --workload $w --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > "$OUT/ncu_launch_$w.log" 2>&1
        echo "ncu launches $w rc=$?"
        # full capture of the pair kernels of one warm evaluation (4th call)
        timeout 600 ncu --set full --clock-control none --import-source on \
            -k 'regex:k_forces|k_neighbours|k_density|k_correction|g_walk' -s 9 -c 3 -f -o "$OUT/full_$w" \
            python bench.py --workload $w --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > "$OUT/ncu_full_$w.log" 2>&1
        echo "ncu full $w rc=$?"
    done
fi

if has drift; then
    timeout 600 python tools/drift_check.py --configs sedov --out "$OUT/drift_sedov.json" > "$OUT/drift.log" 2>&1
    echo "drift rc=$?"
    tail -n 3 "$OUT/drift.log"
fi
ls -la "$OUT"
