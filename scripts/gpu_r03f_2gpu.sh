#!/bin/bash
# 2-GPU box: where do the +20 ms of a DDP training step come from?  (135.5 vs 115.5 ms)
mkdir -p gpurun_out
python -m prior_flow_b200.build > gpurun_out/build.log 2>&1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/train_bench.py --steps 6 --warmup 3"
for mode in none ddp static nobroadcast; do
  timeout 300 $T --ddp $mode 2>/dev/null | tail -1 | tee -a gpurun_out/r03f_train_2gpu.jsonl
done
timeout 300 python scripts/train_bench.py --steps 6 --warmup 3 2>/dev/null | tail -1 | tee -a gpurun_out/r03f_train_2gpu.jsonl
# 2-GPU DDP step == 1-GPU step with the same global batch
timeout 600 python -m pytest tests/test_gpu_ddp.py -m gpu -q -s --timeout 500 -p no:cacheprovider 2>&1 | tail -5
