#!/bin/bash
# Final single-GPU call of a round: whole GPU suite, smoke, the committed profile set of the shipped build, bench (both arms).
#   profiles/<tag>_layers.csv / _traffic.json / _conv_full.csv come from scripts/summarize_ncu.py on the CSVs written here.
set -x
TAG=${1:-final}
O=gpurun_out/$TAG
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q -s > $O/pytest_gpu.log 2>&1
timeout 600 python __graft_entry__.py smoke > $O/smoke.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none --csv --log-file $O/launches_metrics.csv python scripts/profile_step.py > $O/ncu_list.log 2>&1
PAGES=32 timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:conv_tc -c 40 -o /tmp/conv_full python scripts/profile_step.py > $O/ncu_full.log 2>&1
ncu -i /tmp/conv_full.ncu-rep --page raw --csv > $O/conv_full_raw.csv 2>> $O/ncu_full.log
timeout 300 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
du -sh gpurun_out
grep -E "^\.?F?C[234]:|jpeg |passed|failed" $O/pytest_gpu.log | cut -c1-300
tail -n 3 $O/smoke.log
cat $O/bench.json $O/bench_reference.json
