#!/bin/bash
# 8-GPU box: BASELINE configs[2] (global batch 64, strong) and [4] (train step, batch 8 on 8 GPUs, DDP) + the weak-scaling line.
mkdir -p gpurun_out
python -m prior_flow_b200.build > gpurun_out/build.log 2>&1
N=${1:-8}
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
timeout 600 $R bench.py --gpus $N --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r03h_bench_${N}gpu_weak.json; cut -c1-330 gpurun_out/r03h_bench_${N}gpu_weak.json
timeout 900 $R bench.py --gpus $N --global-batch 64 --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r03h_bench_${N}gpu_batch64_strong.json; cut -c1-330 gpurun_out/r03h_bench_${N}gpu_batch64_strong.json
timeout 600 $R scripts/train_bench.py --batch 1 --steps 6 --warmup 3 2>/dev/null | tail -1 | tee gpurun_out/r03h_train_${N}gpu_batch${N}.json
timeout 600 python -m pytest tests/test_gpu_ddp.py -m gpu -q -s --timeout 500 -p no:cacheprovider 2>&1 | grep -E "\[ddp\]|passed|failed"
