#!/bin/bash
# Round-2 seventh GPU call: width of the single A box (16 vs 10 pixels; every grouped layer vs the N <= 64 rule), CRNN shapes in the
# MMA wait-counter probe, the remaining new tests (colour, stn=False, caller quads, nvJPEG).
set -x
O=gpurun_out/r2g
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q > $O/pytest_gpu.log 2>&1
for cfg in "16 0" "10 0" "10 1" "16 1"; do
  set -- $cfg
  B2O_TC_BOX16=$1 B2O_TC_BOX_ALL=$2 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_bytes.sum --clock-control none --csv --log-file $O/launches_box$1_all$2.csv python scripts/profile_step.py > $O/ncu_box$1_all$2.log 2>&1
done
B2O_LIB=$PWD/keras-ocr_b200/libb2ocr_dbg.so timeout 300 python scripts/dev_tc_debug.py > $O/tc_debug.log 2>&1
for f in $O/*.log; do echo "== $f"; tail -n 14 $f; done
