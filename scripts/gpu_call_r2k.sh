#!/bin/bash
set -x
O=gpurun_out/r2k
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "variants or craft or golden or conv" > $O/pytest.log 2>&1
timeout 400 python scripts/dev_pair_ab.py B2O_TC_AFF smem const > $O/ab_aff.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file $O/launches_metrics.csv python scripts/profile_step.py > $O/ncu_list.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err
tail -3 $O/pytest.log; cat $O/ab_aff.log; cat $O/bench.json | cut -c1-160
