#!/bin/bash
# Runs on the GPU box (under gpurun): GPU parity tests, then the per-kernel micro-benchmark.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu_info.csv 2>&1
python -m prior_flow_b200.build > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider "$@" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
timeout 600 python scripts/kbench.py --iters 10 > gpurun_out/kbench.log 2>&1
tail -20 gpurun_out/kbench.log
