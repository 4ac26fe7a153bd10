#!/bin/bash
# Light validation call (when the GPU budget is short): whole GPU suite, smoke, bench, then the launch list of one step.
set -x
TAG=${1:-light}
O=gpurun_out/$TAG
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -s -x > $O/pytest_gpu.log 2>&1
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none --csv --log-file $O/launches_metrics.csv python scripts/profile_step.py > $O/ncu_list.log 2>&1
grep -E "^\.?F?C[234]:|passed|failed|Error" $O/pytest_gpu.log | cut -c1-300
tail -n 2 $O/smoke.log
cut -c1-400 $O/bench.json
grep -E "quads|select_kernel" $O/launches_metrics.csv | grep duration | cut -d, -f5,15
