#!/bin/bash
# Round 2, call F: drop-in binary (rk2_adaptive + monaghan_pc) and device-side conserved quantities.
set -u
OUT=gpurun_out/${1:-r2f}
mkdir -p "$OUT"
timeout 1800 python -m pytest tests/test_gpu_dropin_and_conserved.py -m gpu -q -p no:cacheprovider --durations=8 > "$OUT/pytest.log" 2>&1
echo "pytest rc=$?"; grep -n "^E  \|passed\|failed" "$OUT/pytest.log" | head -40
cat gpurun_out/dropin/*.json 2>/dev/null
