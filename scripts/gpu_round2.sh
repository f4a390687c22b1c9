#!/bin/bash
mkdir -p gpurun_out
python -m prior_flow_b200.build > gpurun_out/build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python scripts/kbench.py --iters 10 > gpurun_out/kbench.log 2>&1; tail -12 gpurun_out/kbench.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -5 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
