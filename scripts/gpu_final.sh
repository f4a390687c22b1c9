#!/bin/bash
# GPU box: the round's closing evidence — full GPU tests, smoke, bench (both arms), ncu launch list + full captures.
mkdir -p gpurun_out
python -m prior_flow_b200.build > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 300 python scripts/kbench.py --iters 20 > gpurun_out/kbench.log 2>&1
timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -1 gpurun_out/bench.log | cut -c1-400; tail -3 gpurun_out/bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2> gpurun_out/bench_reference.err; tail -1 gpurun_out/bench_reference.log | cut -c1-300
PF_TAG=final_cl PF_CHANNELS_LAST=1 PF_CUDNN_BENCHMARK=1 timeout 300 python scripts/e2e_breakdown.py > gpurun_out/e2e.log 2>&1
PF_TAG=final_nchw PF_CUDNN_BENCHMARK=1 timeout 300 python scripts/e2e_breakdown.py >> gpurun_out/e2e.log 2>&1
bash scripts/profile.sh lookup_rows_kernel rotate_fwd_kernel volume_tc_kernel > gpurun_out/profile.log 2>&1
# launch list of the bench command itself (eager launches, 1 timed step), as the profiling recipe asks
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ls -la gpurun_out | head -40
