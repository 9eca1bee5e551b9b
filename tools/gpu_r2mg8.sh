#!/bin/bash
# Round 2, 8-GPU call: impact 8M (1M per GPU, what the driver's scaling run does) and 16M (BASELINE.json's multi-GPU config).
set -u
OUT=gpurun_out/${1:-r2mg8}
mkdir -p "$OUT"
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > "$OUT/gpu.txt" 2>&1
free -g > "$OUT/mem.txt"; nproc >> "$OUT/mem.txt"; cat "$OUT/mem.txt"
run() {  # <tag> <extra args>
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
        bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline $2 > "$OUT/bench_$1.json" 2> "$OUT/bench_$1.err"
    echo "bench $1 rc=$?"; tail -n 3 "$OUT/bench_$1.err" | cut -c1-300
    python - "$OUT/bench_$1.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(" value %.4g ms/step %.3f e2e %s" % (d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value")))
    print(" parity", {k: d["parity"][k] for k in ("max_rel_err", "worst_field", "noi_mismatches", "ok")} if d.get("parity") else None)
    print(" ranks", d["config"]["ranks"]["rows"])
except Exception as e:
    print(" unreadable", e)
PY
}
run impact8M ""
AVAIL=$(free -g | awk '/^Mem:/ {print $7}')
if [ "${AVAIL:-0}" -ge 400 ]; then run impact16M "--particles 2000000 --no-e2e"; else echo "skipping 16M: only ${AVAIL} GB of host memory available"; fi
