#!/bin/bash
# same box, back to back: ncu launch lists with the epilogue constants in shared memory (round 1) and in the constant bank
set -x
O=gpurun_out/r2l
mkdir -p $O
M=gpu__time_duration.sum,lts__t_bytes.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed
B2O_TC_AFF=smem timeout 600 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file $O/launches_smem.csv python scripts/profile_step.py > $O/ncu_smem.log 2>&1
timeout 600 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file $O/launches_const.csv python scripts/profile_step.py > $O/ncu_const.log 2>&1
B2O_TC_AFF=smem timeout 600 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file $O/launches_smem2.csv python scripts/profile_step.py > $O/ncu_smem2.log 2>&1
timeout 600 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file $O/launches_const2.csv python scripts/profile_step.py > $O/ncu_const2.log 2>&1
B2O_TC_AFF=smem timeout 300 python bench.py --steps 10 --warmup 3 > $O/bench_smem.json 2> $O/bench_smem.err
timeout 300 python bench.py --steps 10 --warmup 3 > $O/bench_const.json 2> $O/bench_const.err
B2O_TC_AFF=smem timeout 300 python bench.py --steps 10 --warmup 3 > $O/bench_smem2.json 2> $O/bench_smem2.err
timeout 300 python bench.py --steps 10 --warmup 3 > $O/bench_const2.json 2> $O/bench_const2.err
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "crnn or variants" > $O/pytest.log 2>&1; tail -2 $O/pytest.log
