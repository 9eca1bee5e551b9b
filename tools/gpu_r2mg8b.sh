#!/bin/bash
# Round 2, second 8-GPU call: impact 8M with the C++/NCCL host (default) and the reorder on every rank; 16M if time allows.
set -u
OUT=gpurun_out/${1:-r2mg8b}
mkdir -p "$OUT"
run() {  # <tag> <extra args>
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 \
        bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e $2 > "$OUT/bench_$1.json" 2> "$OUT/bench_$1.err"
    echo "bench $1 rc=$?"; grep -v "Warning\|OMP_NUM\|\*\*\*\*" "$OUT/bench_$1.err" | tail -n 3 | cut -c1-300
    python - "$OUT/bench_$1.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(" value %.4g ms/step %.3f host %s" % (d["value"], d["ms_per_step"], d["config"].get("multi_gpu_host")))
    print(" parity", {k: d["parity"][k] for k in ("max_rel_err", "worst_field", "noi_mismatches", "ok")} if d.get("parity") else None)
    print(" ranks", d["config"]["ranks"]["rows"])
except Exception as e:
    print(" unreadable", e)
PY
}
run impact8M_native ""
run impact16M_native "--particles 2000000"
