#!/bin/bash
# Retry wrapper around gpurun: exit code 3 = no slot free right now (nothing charged) -> wait and retry.
#   scripts/gpu.sh 600 'python -m pytest tests -m gpu -x -q'
t=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 150
done
exit 3
