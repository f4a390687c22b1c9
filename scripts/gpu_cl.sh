#!/bin/bash
mkdir -p gpurun_out
python -m prior_flow_b200.build > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_e2e.py -m gpu -q --timeout 600 -p no:cacheprovider -x 2>&1 | tail -5
timeout 600 python bench.py --skip-cpu-baseline 2> gpurun_out/bench_cl.err | tail -1 | cut -c1-300
timeout 600 python bench.py --skip-cpu-baseline --memory-format nchw 2> gpurun_out/bench_nchw.err | tail -1 | cut -c1-300
