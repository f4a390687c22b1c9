#!/bin/bash
# GPU box: ncu launch list of the kernel micro-benchmark + one full capture per hot kernel ($@ = kernel name regexes).
mkdir -p gpurun_out
python -m prior_flow_b200.build > gpurun_out/build.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_kbench.csv \
    python scripts/kbench.py --iters 2 --skip-torch > gpurun_out/ncu_kbench.log 2>&1
for k in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof_$k \
      python scripts/kbench.py --iters 1 --skip-torch > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out
