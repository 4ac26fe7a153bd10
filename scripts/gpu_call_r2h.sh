#!/bin/bash
# Round-2 eighth GPU call: the whole GPU suite (BASELINE-size parity with decisive weights, colour, nvJPEG, ...), smoke, bench.
set -x
O=gpurun_out/r2h
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q -s > $O/pytest_gpu.log 2>&1
timeout 600 python __graft_entry__.py smoke > $O/smoke.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err
for f in $O/*.log; do echo "== $f"; tail -n 25 $f; done
cat $O/bench.json
