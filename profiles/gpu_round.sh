#!/bin/bash
# One gpurun call: GPU test files in separate processes (a poisoned CUDA context does not take the
# other files down), then the bench lines.  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
export CLIPB200_SYNTHETIC_WEIGHTS=1
rc_all=0
for f in "$@"; do
  name=$(basename "$f" .py)
  timeout 900 python -m pytest "$f" -m gpu -q --timeout 600 -x > "gpurun_out/t_${name}.log" 2>&1
  rc=$?
  echo "== $f rc=$rc: $(tail -1 gpurun_out/t_${name}.log)"
  if [ $rc -ne 0 ]; then rc_all=1; grep -E "^(FAILED|ERROR)|Error|error:|assert" "gpurun_out/t_${name}.log" | head -12; fi
done
exit $rc_all
