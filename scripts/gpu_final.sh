#!/bin/bash
# GPU box: the round's closing evidence — full GPU tests, smoke, clean kernel timings, bench (both arms, both conv precisions), per-kernel
# breakdown of a forward, ncu launch list of the bench command + full captures of the hot kernels, sanitizer on the new kernels.
TAG=${1:-r03z}
mkdir -p gpurun_out
python -m prior_flow_b200.build > gpurun_out/build.log 2>&1
python scripts/vendor_reference.py --check > gpurun_out/vendor_check.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s --timeout 900 -p no:cacheprovider > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit: $?" >> gpurun_out/${TAG}_pytest_gpu.log
grep -E "passed|failed" gpurun_out/${TAG}_pytest_gpu.log | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
# kernel timings: a clean run (NOT under ncu) is the only thing that writes the JSON
timeout 600 python scripts/kbench.py --iters 20 --out gpurun_out/${TAG}_kbench.json > gpurun_out/${TAG}_kbench.log 2>&1; grep -c kernel gpurun_out/${TAG}_kbench.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench_line.json 2> gpurun_out/bench.err; cut -c1-400 gpurun_out/${TAG}_bench_line.json; tail -2 gpurun_out/bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_line.json 2> gpurun_out/bench_reference.err; cut -c1-300 gpurun_out/${TAG}_bench_reference_line.json
timeout 900 python bench.py --cudnn-tf32 off --skip-cpu-baseline --skip-gpu-baselines --skip-traffic > gpurun_out/${TAG}_bench_line_fp32_convs.json 2>/dev/null; cut -c1-200 gpurun_out/${TAG}_bench_line_fp32_convs.json
PF_TAG=${TAG}_cl PF_CHANNELS_LAST=1 PF_CUDNN_BENCHMARK=1 timeout 300 python scripts/e2e_breakdown.py > gpurun_out/e2e.log 2>&1
# launch list of the bench command itself (eager launches, 1 timed step), as the profiling recipe asks
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu-baseline --skip-gpu-baselines --skip-traffic > gpurun_out/ncu_bench.log 2>&1
for k in otf_dots_kernel otf_blend_kernel otf_box_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/${TAG}_$k python scripts/probe/otf_tiles.py > gpurun_out/ncu_$k.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"otf_|rotate" --csv --log-file gpurun_out/${TAG}_launches_onthefly_tc.csv python scripts/probe/otf_tiles.py > gpurun_out/${TAG}_onthefly_tile_boxes.txt 2>&1
timeout 300 python scripts/probe/otf_hires.py > gpurun_out/${TAG}_onthefly_hires_call.txt 2>&1; tail -1 gpurun_out/${TAG}_onthefly_hires_call.txt
for k in lookup_rows_kernel rotate_fwd_kernel volume_tc_kernel dccl_conv_kernel volume_bwd_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/${TAG}_$k \
      python scripts/kbench.py --iters 1 --skip-torch --only "lookup_dual,lookup_conv[fp32,volume_pyramid[fp32],volume_backward[tcgen05" > gpurun_out/ncu_$k.log 2>&1
done
# compute-sanitizer over the new kernels' small-shape tests
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 --target-processes all \
  python -m pytest tests/test_gpu_conv.py tests/test_gpu_sphere.py tests/test_gpu_configs.py tests/test_gpu_onthefly_tc.py -q -p no:cacheprovider -x --timeout 800 \
  -k "2-16-32 or 24-44 or convex_upsample or uniform_loss or great_circle or grad_sink or (volume_backward and 16-32) or 1-16-32 or single_view" > gpurun_out/${TAG}_sanitizer_memcheck.log 2>&1
echo "memcheck exit: $?" >> gpurun_out/${TAG}_sanitizer_memcheck.log; tail -3 gpurun_out/${TAG}_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 99 --target-processes all \
  python -m pytest tests/test_gpu_conv.py tests/test_gpu_sphere.py tests/test_gpu_onthefly_tc.py -q -p no:cacheprovider -x --timeout 800 -k "2-16-32 or convex_upsample or uniform_loss or smooth-1-16-32" > gpurun_out/${TAG}_sanitizer_racecheck.log 2>&1
echo "racecheck exit: $?" >> gpurun_out/${TAG}_sanitizer_racecheck.log; tail -3 gpurun_out/${TAG}_sanitizer_racecheck.log
# the other BASELINE configs for the record (configs[3]: 1024x2048 / 32 iterations, on-the-fly as named and materialised as `auto` picks)
X="--skip-cpu-baseline --skip-gpu-baselines --skip-traffic"
timeout 900 python bench.py $X --height 1024 --width 2048 --iters 32 --steps 3 --corr-mode onthefly 2> gpurun_out/hi1.err | tail -1 > gpurun_out/${TAG}_bench_hires_1024x2048_onthefly.json; cut -c1-200 gpurun_out/${TAG}_bench_hires_1024x2048_onthefly.json
PF_ONTHEFLY_TC=0 timeout 900 python bench.py $X --height 1024 --width 2048 --iters 32 --steps 3 --corr-mode onthefly 2> gpurun_out/hi3.err | tail -1 > gpurun_out/${TAG}_bench_hires_1024x2048_onthefly_cuda_cores.json; cut -c1-200 gpurun_out/${TAG}_bench_hires_1024x2048_onthefly_cuda_cores.json
timeout 900 python bench.py $X --height 1024 --width 2048 --iters 32 --steps 3 --corr-mode materialized 2> gpurun_out/hi2.err | tail -1 > gpurun_out/${TAG}_bench_hires_1024x2048_materialized.json; cut -c1-200 gpurun_out/${TAG}_bench_hires_1024x2048_materialized.json
ls gpurun_out | grep ${TAG} | head -60
