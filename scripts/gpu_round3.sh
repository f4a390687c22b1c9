#!/bin/bash
mkdir -p gpurun_out
python -m prior_flow_b200.build > gpurun_out/build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_gpu.log
grep -E "^E  +(Assertion|assert)|passed|failed|^FAILED|Error" gpurun_out/pytest_gpu.log | head -20
timeout 300 python scripts/kbench.py --iters 10 --skip-torch > gpurun_out/kbench.log 2>&1; grep -E "lookup|volume_pyramid\[fp32\]" gpurun_out/kbench.log
PF_TAG=cl_bench PF_CHANNELS_LAST=1 PF_CUDNN_BENCHMARK=1 timeout 300 python scripts/e2e_breakdown.py > gpurun_out/e2e_cl.log 2>&1; head -16 gpurun_out/e2e_breakdown_cl_bench.txt | cut -c1-150
for mf in nchw channels_last; do timeout 600 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --memory-format $mf > gpurun_out/bench_$mf.log 2> gpurun_out/bench_$mf.err; python -c "
import json,sys
l=json.loads(open('gpurun_out/bench_$mf.log').read().strip().splitlines()[-1]); print('$mf', l['value'], l['ms_per_step'], 'e2e', l['e2e']['value'], l['roofline']['ms_per_launch'], l['roofline']['frac'], l['gpu_launches_per_step'])" ; tail -2 gpurun_out/bench_$mf.err; done
