#!/bin/bash
# GPU box, r03c: full GPU tests + bench with the new legs (reference on CPU / on the GPU, drop-in, measured traffic, gather floor).
mkdir -p gpurun_out
python -m prior_flow_b200.build > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_gpu.log
grep -E "^\[|passed|failed|rror" gpurun_out/pytest_gpu.log | tail -20
timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -1 gpurun_out/bench.log | cut -c1-3000; tail -5 gpurun_out/bench.err
