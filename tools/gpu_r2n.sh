#!/bin/bash
# Round 2, call N (2 GPUs): the Python multi-GPU host once more after the stream-binding change in HaloExchange.
set -u
OUT=gpurun_out/${1:-r2n}
mkdir -p "$OUT"
timeout 200 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -x -p no:cacheprovider > "$OUT/pytest_mg.log" 2>&1
echo "pytest mg rc=$?"; tail -n 5 "$OUT/pytest_mg.log"
