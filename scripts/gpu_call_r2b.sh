#!/bin/bash
# Round-2 second GPU call: single-box tiles with the corrected descriptor (default on), fused tail, liveness-packed workspace.
set -x
O=gpurun_out/r2b
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1
timeout 400 python scripts/dev_pair_ab.py B2O_TC_BOX16 0 1 > $O/ab_box16.log 2>&1
timeout 400 python scripts/dev_pair_ab.py B2O_FUSED_TAIL 0 1 > $O/ab_tail.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none --csv --log-file $O/launches_metrics.csv python scripts/profile_step.py > $O/ncu_list.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err
du -sh gpurun_out
for f in $O/*.log; do echo "== $f"; tail -n 8 $f; done
cat $O/bench.json
