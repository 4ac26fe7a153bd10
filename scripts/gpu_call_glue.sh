#!/bin/bash
# One `--set full` capture of the step's non-conv kernels -> <tag>/glue_full_raw.csv
set -x
TAG=${1:-glue}
O=gpurun_out/$TAG
mkdir -p $O
PAGES=32 timeout 900 ncu --profile-from-start off --set full --clock-control none \
  -k 'regex:upsample|quads|maxpool|stn_|normalize16|resize_pad|lstm|fc_ctc|select_k|merge_k|flatten_k|binarize|stats_k|warp_kernel|stem_crnn|add_kernel|compact|crops_to|pack_rec|gray' \
  -c 80 -o /tmp/glue_full python scripts/profile_step.py > $O/ncu_glue.log 2>&1
ncu -i /tmp/glue_full.ncu-rep --page raw --csv > $O/glue_full_raw.csv 2>> $O/ncu_glue.log
tail -3 $O/ncu_glue.log; wc -l $O/glue_full_raw.csv
