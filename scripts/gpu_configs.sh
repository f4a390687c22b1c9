#!/bin/bash
# GPU box: the other BASELINE configs for the record (not bench lines): batch 8 per GPU, 1024x2048 / 32 iters on-the-fly and materialised.
mkdir -p gpurun_out
python -m prior_flow_b200.build > gpurun_out/build.log 2>&1
timeout 600 python bench.py --skip-cpu-baseline --batch 8 --steps 5 2> gpurun_out/b8.err | tail -1 > gpurun_out/bench_batch8.json; cut -c1-330 gpurun_out/bench_batch8.json; tail -2 gpurun_out/b8.err
timeout 900 python bench.py --skip-cpu-baseline --height 1024 --width 2048 --iters 32 --steps 3 --corr-mode onthefly 2> gpurun_out/hi1.err | tail -1 > gpurun_out/bench_hires_onthefly.json; cut -c1-330 gpurun_out/bench_hires_onthefly.json; tail -2 gpurun_out/hi1.err
timeout 900 python bench.py --skip-cpu-baseline --height 1024 --width 2048 --iters 32 --steps 3 --corr-mode materialized 2> gpurun_out/hi2.err | tail -1 > gpurun_out/bench_hires_materialized.json; cut -c1-330 gpurun_out/bench_hires_materialized.json; tail -2 gpurun_out/hi2.err
