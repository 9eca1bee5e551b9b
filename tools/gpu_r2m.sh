#!/bin/bash
# Round 2, call M: the driver's default line once more on the final tree (roofline.traffic from the round-2 capture,
# l1tex_data_pipe beside the FP64 fraction), plus sedov.
set -u
OUT=gpurun_out/${1:-r2m}
mkdir -p "$OUT"
timeout 400 python bench.py > "$OUT/bench_impact.json" 2> "$OUT/bench_impact.err"; echo "impact rc=$?"
timeout 400 python bench.py --workload sedov --no-cpu-baseline > "$OUT/bench_sedov.json" 2> "$OUT/bench_sedov.err"; echo "sedov rc=$?"
python tools/show_bench.py "$OUT/bench_impact.json" "$OUT/bench_sedov.json"
python -c "
import json
d=json.loads(open('$OUT/bench_impact.json').read().strip().splitlines()[-1]); r=d['roofline']; print(r['traffic'], r['traffic_source'], r['l1tex_data_pipe'])"
