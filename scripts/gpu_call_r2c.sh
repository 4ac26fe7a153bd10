#!/bin/bash
# Round-2 third GPU call: where does the stem's time go (MMA-warp wait counters of the debug build, with and without epilogue stores),
# the per-layer single-box rule, bench.
set -x
O=gpurun_out/r2c
mkdir -p $O
B2O_LIB=$PWD/keras-ocr_b200/libb2ocr_dbg.so timeout 300 python scripts/dev_tc_debug.py > $O/tc_debug.log 2>&1
B2O_DEBUG_NOSTORE=1 B2O_LIB=$PWD/keras-ocr_b200/libb2ocr_dbg.so timeout 300 python scripts/dev_tc_debug.py > $O/tc_debug_nostore.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -k "conv or craft or bit" > $O/pytest_conv.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none --csv --log-file $O/launches_metrics.csv python scripts/profile_step.py > $O/ncu_list.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err
for f in $O/*.log; do echo "== $f"; tail -n 12 $f; done
cat $O/bench.json
