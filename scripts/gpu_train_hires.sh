#!/bin/bash
# train step at 512x1024 and 1024x2048, materialised vs volume-free (on-the-fly) correlation: time and peak memory
mkdir -p gpurun_out
: > gpurun_out/r03z_train_step_corr_modes.jsonl
for shape in "512 1024 12" "1024 2048 12"; do
  set -- $shape
  for mode in materialized onthefly; do
    timeout 600 python scripts/train_bench.py --height $1 --width $2 --iters $3 --corr-mode $mode --steps 3 --warmup 2 2> gpurun_out/train_$mode.err | tail -1 >> gpurun_out/r03z_train_step_corr_modes.jsonl
    tail -1 gpurun_out/train_$mode.err | cut -c1-200
  done
done
cut -c1-260 gpurun_out/r03z_train_step_corr_modes.jsonl
