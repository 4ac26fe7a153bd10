#!/bin/bash
# Source-level (PC sampling) view of quads_kernel in one profiled step -> <tag>/quads_source.csv
set -x
TAG=${1:-quads}
O=gpurun_out/$TAG
mkdir -p $O
PAGES=32 timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none \
  -k 'regex:quads_kernel' -c 1 -o /tmp/quads python scripts/profile_step.py > $O/ncu.log 2>&1
ncu -i /tmp/quads.ncu-rep --page source --csv > $O/quads_source.csv 2>> $O/ncu.log
ncu -i /tmp/quads.ncu-rep --page raw --csv > $O/quads_raw.csv 2>> $O/ncu.log
tail -2 $O/ncu.log; wc -l $O/quads_source.csv
