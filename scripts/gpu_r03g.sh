#!/bin/bash
# GPU box, r03g: full GPU test suite + kernel timings of the new kernels + 1-GPU train step + default bench.
mkdir -p gpurun_out
python -m prior_flow_b200.build > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_gpu.log
grep -E "^\[|passed|failed" gpurun_out/pytest_gpu.log | tail -30
timeout 300 python scripts/kbench.py --iters 20 --only volume_backward,convex_upsample --out gpurun_out/kbench_r03g.json 2>&1 | grep kernel
timeout 300 python scripts/train_bench.py --steps 6 --warmup 3 2>/dev/null | tail -1 | tee gpurun_out/r03g_train_1gpu.json
timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -1 gpurun_out/bench.log | cut -c1-1200; tail -3 gpurun_out/bench.err
