#!/bin/bash
# Round-2 fifth GPU call: commuted decoder upsampling (UPADD epilogue) -- parity, A/B, launch list, bench.
set -x
O=gpurun_out/r2e
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/pytest_gpu.log 2>&1
timeout 400 python scripts/dev_pair_ab.py B2O_UPCONV_COMMUTE 0 1 > $O/ab_commute.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none --csv --log-file $O/launches_metrics.csv python scripts/profile_step.py > $O/ncu_list.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err
for f in $O/*.log; do echo "== $f"; tail -n 12 $f; done
cat $O/bench.json
