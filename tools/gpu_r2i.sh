#!/bin/bash
# Round 2, call I: the self-gravitating workloads (step-0 state; the reference integrator cannot advance them at 10^6
# particles within the bench's time limit) incl. the reference arm, and the evolve-timeout fallback of bench.py.
set -u
OUT=gpurun_out/${1:-r2i}
mkdir -p "$OUT"
timeout 400 python bench.py --workload giant_hydro --state step0 --no-cpu-baseline > "$OUT/bench_giant_hydro.json" 2> "$OUT/bench_giant_hydro.err"; echo "giant_hydro rc=$?"
timeout 400 python bench.py --workload giant_solid --state step0 --no-cpu-baseline > "$OUT/bench_giant_solid.json" 2> "$OUT/bench_giant_solid.err"; echo "giant_solid rc=$?"
timeout 400 python bench.py --impl reference --workload giant_hydro --state step0 --steps 5 --warmup 2 > "$OUT/bench_reference_giant_hydro.json" 2> "$OUT/bench_reference_giant_hydro.err"; echo "ref giant_hydro rc=$?"
python tools/show_bench.py "$OUT/bench_giant_hydro.json" "$OUT/bench_giant_solid.json"
tail -c 600 "$OUT/bench_reference_giant_hydro.json"; echo
# fallback path: a 15 s limit on the reference integrator -> both arms must report the step-0 state and say why
B200SPH_EVOLVE_TIMEOUT_S=15 timeout 400 python bench.py --workload giant_hydro --particles 300000 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > "$OUT/bench_fallback.json" 2> "$OUT/bench_fallback.err"; echo "fallback rc=$?"
python -c "
import json,sys
d=json.loads(open('$OUT/bench_fallback.json').read().strip().splitlines()[-1]); print(d['config']['state'], '|', d['config']['state_note'])"
tail -n 3 "$OUT"/*.err
