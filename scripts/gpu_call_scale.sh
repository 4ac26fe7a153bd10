#!/bin/bash
# Scaling check on ONE multi-GPU box, as the driver does it: N = 1, then N = all, back to back (weak scaling, 32 pages per GPU).
set -x
N=${1:-8}
O=gpurun_out/scale$N
mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.max.sm,power.limit --format=csv > $O/smi.txt
timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 > $O/bench_1gpu.json 2> $O/bench_1gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_${N}gpu.json 2> $O/bench_${N}gpu.err
B2O_BENCH_STREAM=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_${N}gpu_gather.json 2> $O/bench_${N}gpu_gather.err
python - <<PY
import json
a=json.loads(open("$O/bench_1gpu.json").read().strip().splitlines()[-1])
for tag in ("${N}gpu","${N}gpu_gather"):
    b=json.loads(open("$O/bench_"+tag+".json").read().strip().splitlines()[-1])
    print(tag, "value", round(b["value"],1), "e2e", round(b["e2e"]["value"],1), "efficiency vs N=1 on this box", round(b["value"]/($N*a["value"]),4),
          "per-rank ms", b.get("per_rank_ms_per_step"), "decode ms", b.get("rank0_decode_ms_per_step"), "words", b.get("words_per_step"), b.get("words_read_correctly"))
print("1gpu", round(a["value"],1), a["clocks"])
PY
tail -n 3 $O/*.err
