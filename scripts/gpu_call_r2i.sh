#!/bin/bash
# Round-2 ninth GPU call: BASELINE-size parity diagnostics (restructured assertions), compute-sanitizer over a small end-to-end step.
set -x
O=gpurun_out/r2i
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_baseline_sizes.py tests/test_gpu_parity.py -m gpu -q -s -k "c3 or c4 or jpeg or c2" > $O/pytest_sizes.log 2>&1
bash scripts/gpu_sanitizer.sh $O > $O/sanitizer_run.log 2>&1
grep -E "^C[234]:|^jpeg|passed|failed" $O/pytest_sizes.log
cat $O/summary.txt
