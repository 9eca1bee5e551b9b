#!/bin/bash
# Round 2, multi-GPU call: NCCL tests + bench at N GPUs (run under `gpurun --gpus N`).  MEASUREMENT infrastructure.
set -u
N=${1:-2}
OUT=gpurun_out/${2:-r2mg$N}
WL=${3:-impact sedov giant_hydro}
mkdir -p "$OUT"
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > "$OUT/gpu.txt" 2>&1
if [ "$N" = 2 ]; then
timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -p no:cacheprovider > "$OUT/pytest_mg.log" 2>&1
echo "pytest mg rc=$?"; grep -n "^E  \|passed\|failed\|MISMATCH\|EXCEPTION" "$OUT/pytest_mg.log" | head -30
fi
for w in $WL; do
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $N --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > "$OUT/bench_${w}_$N.json" 2> "$OUT/bench_${w}_$N.err"
    echo "bench $w x$N rc=$?"; tail -n 4 "$OUT/bench_${w}_$N.err" | cut -c1-300
    python - "$OUT/bench_${w}_$N.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(" value %.4g ms/step %.3f parity %s" % (d["value"], d["ms_per_step"], d.get("parity")))
    print(" ranks", d["config"]["ranks"]["rows"])
    print(" ", d["config"]["multi_gpu"][:400])
except Exception as e:
    print(" unreadable", e)
PY
done
