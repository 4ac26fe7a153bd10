#!/bin/bash
set -x
O=gpurun_out/r2j
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "variants or craft or golden" > $O/pytest.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none --csv --log-file $O/launches_metrics.csv python scripts/profile_step.py > $O/ncu_list.log 2>&1
tail -3 $O/pytest.log
