#!/bin/bash
# GPU box, r03a: full GPU tests (incl. the drop-in tests on the vendored reference) + source-level ncu capture of the lookup kernels.
mkdir -p gpurun_out
python -m prior_flow_b200.build > gpurun_out/build.log 2>&1
python scripts/vendor_reference.py --check > gpurun_out/vendor_check.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_gpu.log
grep -E "^\[|passed|failed|error" gpurun_out/pytest_gpu.log | tail -20
timeout 300 python scripts/kbench.py --iters 20 --skip-torch --only lookup,volume_pyramid,hbm --out gpurun_out/kbench_r03a.json > gpurun_out/kbench.log 2>&1
tail -8 gpurun_out/kbench.log
for k in lookup_rows_kernel rotate_fwd_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/r03a_$k \
      python scripts/kbench.py --iters 1 --skip-torch --only lookup_dual > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out | head -30
