#!/bin/bash
# Round-2 sixth GPU call: per-layer effect of CTA pairs on the generic tiles (B2O_TC_PAIR=2), colour / stn=False / caller-quad tests.
set -x
O=gpurun_out/r2f
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/pytest_gpu.log 2>&1
B2O_TC_PAIR=2 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,lts__t_bytes.sum --clock-control none --csv --log-file $O/launches_pair2.csv python scripts/profile_step.py > $O/ncu_pair2.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,lts__t_bytes.sum --clock-control none --csv --log-file $O/launches_default.csv python scripts/profile_step.py > $O/ncu_default.log 2>&1
for f in $O/*.log; do echo "== $f"; tail -n 12 $f; done
