#!/bin/bash
# Round 2, call J (2 GPUs): NCCL tests incl. the distributed integrator, and the 2-GPU bench line on the final code.
set -u
OUT=gpurun_out/${1:-r2j}
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_multigpu_native.py tests/test_multigpu_gpu.py -m gpu -q -p no:cacheprovider --durations=6 > "$OUT/pytest_mg.log" 2>&1
echo "pytest mg rc=$?"; grep -n "^E  \|passed\|failed\|MISMATCH\|EXCEPTION\|Error" "$OUT/pytest_mg.log" | head -30
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 3 > "$OUT/bench_impact_2.json" 2> "$OUT/bench_impact_2.err"
echo "bench impact x2 rc=$?"; tail -n 3 "$OUT/bench_impact_2.err" | cut -c1-300
python - "$OUT/bench_impact_2.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(" value %.4g ms/step %.3f e2e %s parity %s" % (d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value"), d.get("parity")))
    print(" ranks", d["config"]["ranks"]["rows"])
except Exception as e:
    print(" unreadable", e)
PY
