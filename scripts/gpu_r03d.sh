#!/bin/bash
# GPU box, r03d: timing of the fused conv path (kernel level and end to end) + ncu capture of dccl_conv_kernel.
mkdir -p gpurun_out
python -m prior_flow_b200.build > gpurun_out/build.log 2>&1
timeout 300 python scripts/kbench.py --iters 30 --skip-torch --only lookup --out gpurun_out/kbench_r03d.json 2>&1 | grep kernel
for f in 1 0; do for tf in on off; do
  echo "== PF_FUSE_CONV1=$f cudnn-tf32=$tf"; PF_FUSE_CONV1=$f timeout 600 python bench.py --skip-cpu-baseline --skip-gpu-baselines --skip-traffic --cudnn-tf32 $tf 2>/dev/null | python -c "import sys,json; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(l['value'], l['ms_per_step'], l['e2e']['value'], l['gpu_launches_per_step'])"
done; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dccl_conv_kernel -s 2 -c 1 -f -o gpurun_out/r03d_dccl_conv_kernel \
    python scripts/kbench.py --iters 1 --skip-torch --only "lookup_conv[fp32" > gpurun_out/ncu_conv.log 2>&1
ls -la gpurun_out | grep r03d
