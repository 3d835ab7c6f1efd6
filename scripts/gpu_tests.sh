#!/bin/bash
# Runs each GPU test file in its own process (a trapped kernel kills its CUDA context, not the others)
# with a hard timeout; logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
rc=0
for f in "$@"; do
  name=$(basename "$f" .py)
  timeout 600 python -m pytest "$f" -m gpu -q -rP --tb=short > "gpurun_out/$name.log" 2>&1
  r=$?
  echo "== $f -> exit $r"
  grep -E "rel err|max abs|passed|failed|Error|error|img/s" "gpurun_out/$name.log" | tail -n 40
  [ $r -ne 0 ] && rc=$r
done
exit $rc
