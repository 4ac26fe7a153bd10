#!/bin/bash
set -x
O=gpurun_out/r2m
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py -m gpu -q -k "crnn or c3 or color or bit_repro or batch" > $O/pytest.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -k regex:stem_crnn --csv --log-file $O/stem_crnn.csv python scripts/profile_step.py > $O/ncu.log 2>&1
tail -2 $O/pytest.log; grep stem_crnn $O/stem_crnn.csv | cut -d, -f5,15
