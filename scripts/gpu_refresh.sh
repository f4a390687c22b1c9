#!/bin/bash
# GPU box: refresh the parts of the closing evidence that the last commits touched (tests, kernel timings, on-the-fly records, bench line)
TAG=${1:-r03z}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s --timeout 900 -p no:cacheprovider > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit: $?" >> gpurun_out/${TAG}_pytest_gpu.log
grep -E "passed|failed" gpurun_out/${TAG}_pytest_gpu.log | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
timeout 600 python scripts/kbench.py --iters 20 --out gpurun_out/${TAG}_kbench.json > gpurun_out/${TAG}_kbench.log 2>&1; grep -c kernel gpurun_out/${TAG}_kbench.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench_line.json 2> gpurun_out/bench.err; cut -c1-300 gpurun_out/${TAG}_bench_line.json; tail -2 gpurun_out/bench.err
for k in otf_dots_kernel otf_blend_kernel otf_box_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/${TAG}_$k python scripts/probe/otf_tiles.py > gpurun_out/ncu_$k.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"otf_|rotate" --csv --log-file gpurun_out/${TAG}_launches_onthefly_tc.csv python scripts/probe/otf_tiles.py > gpurun_out/${TAG}_onthefly_tile_boxes.txt 2>&1
timeout 300 python scripts/probe/otf_hires.py > gpurun_out/${TAG}_onthefly_hires_call.txt 2>&1; tail -1 gpurun_out/${TAG}_onthefly_hires_call.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 --target-processes all \
  python -m pytest tests/test_gpu_conv.py tests/test_gpu_sphere.py tests/test_gpu_configs.py tests/test_gpu_onthefly_tc.py -q -p no:cacheprovider -x --timeout 800 \
  -k "2-16-32 or 24-44 or convex_upsample or uniform_loss or great_circle or grad_sink or (volume_backward and 16-32) or 1-16-32 or single_view" > gpurun_out/${TAG}_sanitizer_memcheck.log 2>&1
echo "memcheck exit: $?" >> gpurun_out/${TAG}_sanitizer_memcheck.log; tail -3 gpurun_out/${TAG}_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 99 --target-processes all \
  python -m pytest tests/test_gpu_conv.py tests/test_gpu_sphere.py tests/test_gpu_onthefly_tc.py -q -p no:cacheprovider -x --timeout 800 -k "2-16-32 or convex_upsample or uniform_loss or smooth-1-16-32" > gpurun_out/${TAG}_sanitizer_racecheck.log 2>&1
echo "racecheck exit: $?" >> gpurun_out/${TAG}_sanitizer_racecheck.log; tail -3 gpurun_out/${TAG}_sanitizer_racecheck.log
X="--skip-cpu-baseline --skip-gpu-baselines --skip-traffic"
timeout 900 python bench.py $X --height 1024 --width 2048 --iters 32 --steps 3 --corr-mode onthefly 2> gpurun_out/hi1.err | tail -1 > gpurun_out/${TAG}_bench_hires_1024x2048_onthefly.json; cut -c1-200 gpurun_out/${TAG}_bench_hires_1024x2048_onthefly.json
