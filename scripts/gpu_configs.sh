#!/bin/bash
# GPU box: the other BASELINE configs for the record: configs[2] at N = 1 (global batch 64 on one GPU), configs[3] (1024x2048 / 32 iters,
# on-the-fly as named and materialised as `auto` picks on a 180 GB part).
mkdir -p gpurun_out
TAG=${1:-r03y}
python -m prior_flow_b200.build > gpurun_out/build.log 2>&1
X="--skip-cpu-baseline --skip-gpu-baselines --skip-traffic"
timeout 900 python bench.py $X --global-batch 64 --steps 5 2> gpurun_out/b64.err | tail -1 > gpurun_out/${TAG}_bench_1gpu_batch64_strong.json; cut -c1-300 gpurun_out/${TAG}_bench_1gpu_batch64_strong.json; tail -2 gpurun_out/b64.err
timeout 900 python bench.py $X --height 1024 --width 2048 --iters 32 --steps 3 --corr-mode onthefly 2> gpurun_out/hi1.err | tail -1 > gpurun_out/${TAG}_bench_hires_1024x2048_onthefly.json; cut -c1-300 gpurun_out/${TAG}_bench_hires_1024x2048_onthefly.json; tail -2 gpurun_out/hi1.err
timeout 900 python bench.py $X --height 1024 --width 2048 --iters 32 --steps 3 --corr-mode materialized 2> gpurun_out/hi2.err | tail -1 > gpurun_out/${TAG}_bench_hires_1024x2048_materialized.json; cut -c1-300 gpurun_out/${TAG}_bench_hires_1024x2048_materialized.json; tail -2 gpurun_out/hi2.err
