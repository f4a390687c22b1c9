#!/bin/bash
# GPU box: does keeping pyramid levels 2-3 L2-resident (evict_last) speed up the lookups INSIDE the model's loop?
mkdir -p gpurun_out
python -m prior_flow_b200.build > gpurun_out/build.log 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_torch_parity.py -m gpu -q -x --timeout 200 -p no:cacheprovider 2>&1 | tail -2
for h in 0 1; do
  echo "== PF_LOOKUP_L2HINT=$h"
  PF_LOOKUP_L2HINT=$h PF_TAG=l2hint$h PF_CHANNELS_LAST=1 PF_CUDNN_BENCHMARK=1 timeout 300 python scripts/e2e_breakdown.py 2>/dev/null | grep -E "total kernel|lookup_rows|rotate_fwd|dccl_conv"
  PF_LOOKUP_L2HINT=$h timeout 200 python scripts/kbench.py --iters 30 --skip-torch --only lookup_dual 2>&1 | grep kernel
  PF_LOOKUP_L2HINT=$h timeout 600 python bench.py --skip-cpu-baseline --skip-gpu-baselines --skip-traffic 2>/dev/null | python -c "import sys,json; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', l['value'], l['ms_per_step'], l['roofline']['ms_per_launch'], l['roofline']['ms_per_launch_fused_sum'])"
done
