#!/bin/bash
# Round-2 first GPU call: validate the prepared experimental conv variants, A/B them, and capture the shipped pair kernel with ncu --set full.
# (.ncu-rep files stay in /tmp on the box: gpurun_out/ is limited to 64 MiB; only CSV exports come back)
set -x
mkdir -p gpurun_out/r2a
O=gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/smi.txt
( cd scripts/probes && nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o dx_shift_probe dx_shift_probe.cu && timeout 120 ./dx_shift_probe ) > $O/dx_shift_probe.log 2>&1
B2O_EXPERIMENTAL=1 timeout 600 python -m pytest tests -m gpu -k experimental -q > $O/experimental.log 2>&1
timeout 400 python scripts/dev_pair_ab.py B2O_TC_BOX16 0 1 > $O/ab_box16.log 2>&1
timeout 400 python scripts/dev_pair_ab.py B2O_TC_PAIR 1 2 > $O/ab_pair2.log 2>&1
PAGES=32 timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:conv_tc -c 40 -o /tmp/conv_pairs python scripts/profile_step.py > $O/ncu_full.log 2>&1
ncu -i /tmp/conv_pairs.ncu-rep --page raw --csv > $O/conv_pairs_raw.csv 2>> $O/ncu_full.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum --clock-control none --csv --log-file $O/launches_metrics.csv python scripts/profile_step.py > $O/ncu_list.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1
du -sh gpurun_out
for f in $O/*.log; do echo "== $f"; tail -n 6 $f; done
