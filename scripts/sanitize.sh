#!/bin/bash
# GPU box: compute-sanitizer (memcheck, then racecheck) over the small-shape parity tests of every kernel.
mkdir -p gpurun_out
python -m prior_flow_b200.build > gpurun_out/build.log 2>&1
SEL="samplers or img_rotate or flo_rotate_bit or samplegrid or golden or other_direction or corrblock or ragged or channels_last or warp_groupcorr or onthefly_equals or volume_pyramid_vs_golden or avg_pool"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 --target-processes all \
  python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -x --timeout 1200 -k "$SEL" > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck exit: $?" >> gpurun_out/sanitizer_memcheck.log
tail -6 gpurun_out/sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 99 --target-processes all \
  python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -x --timeout 1200 -k "golden or channels_last or warp_groupcorr or onthefly_equals" > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck exit: $?" >> gpurun_out/sanitizer_racecheck.log
tail -6 gpurun_out/sanitizer_racecheck.log
