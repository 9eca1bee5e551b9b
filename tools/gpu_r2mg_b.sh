#!/bin/bash
# Round 2, 2-GPU call B: native C++/NCCL host tests, both hosts in the bench, new golden (impact_aneos).
set -u
OUT=gpurun_out/${1:-r2mgb}
mkdir -p "$OUT/golden"
true
echo "golden rc=$?"; tail -n 2 "$OUT/golden.log"
cp "$OUT"/golden/*.npz tests/golden/ 2>/dev/null
true
echo "pytest aneos rc=$?"; grep -n "^E  \|passed\|failed" "$OUT/pytest_aneos.log" | head
timeout 900 python -m pytest tests/test_multigpu_native.py tests/test_multigpu_gpu.py -m gpu -q -p no:cacheprovider > "$OUT/pytest_mg.log" 2>&1
echo "pytest mg rc=$?"; grep -n "^E  \|passed\|failed\|MISMATCH\|EXCEPTION\|Error" "$OUT/pytest_mg.log" | head -30
for host in python native; do
for w in impact sedov; do
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
        bench.py --gpus 2 --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --mg-host $host > "$OUT/bench_${w}_${host}.json" 2> "$OUT/bench_${w}_${host}.err"
    echo "bench $w $host rc=$?"; grep -v "Warning\|OMP_NUM\|\*\*\*\*" "$OUT/bench_${w}_${host}.err" | tail -n 4 | cut -c1-300
    python - "$OUT/bench_${w}_${host}.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(" value %.4g ms/step %.3f" % (d["value"], d["ms_per_step"]))
    print(" parity", {k: d["parity"][k] for k in ("max_rel_err", "worst_field", "noi_mismatches", "ok")} if d.get("parity") else None)
    print(" ranks", d["config"]["ranks"]["rows"])
except Exception as e:
    print(" unreadable", e)
PY
done
done
