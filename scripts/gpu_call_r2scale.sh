#!/bin/bash
set -x
mkdir -p gpurun_out/scale8
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests/test_gpu_baseline_sizes.py -m gpu -q -s > gpurun_out/scale8/pytest_sizes.log 2>&1
grep -E "C[234]:|passed|failed" gpurun_out/scale8/pytest_sizes.log | cut -c1-300
bash scripts/gpu_call_scale.sh 8
