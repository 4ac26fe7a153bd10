#!/bin/bash
# Round-2 fourth GPU call (2 GPUs): per-step gather vs the pipelined host decode (distributed.ShardedStream), per-rank timings.
set -x
O=gpurun_out/r2d
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/smi.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_2gpu_gather.json 2> $O/bench_2gpu_gather.err
B2O_BENCH_STREAM=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_2gpu_stream.json 2> $O/bench_2gpu_stream.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/bench_2gpu_reference.json 2> $O/bench_2gpu_reference.err
cat $O/*.json
tail -n 5 $O/*.err
