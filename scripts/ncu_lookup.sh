#!/bin/bash
# GPU box: one full ncu capture of the lookup + rotate kernels of a DCCL call (bench-like random coords).
mkdir -p gpurun_out
python -m prior_flow_b200.build > gpurun_out/build.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"lookup|rotate" -s 6 -c 2 -f -o gpurun_out/prof_lookup_call \
    python scripts/lookup_tune.py --reps 1 > gpurun_out/ncu_lookup.log 2>&1
ncu -i gpurun_out/prof_lookup_call.ncu-rep --page raw --csv > gpurun_out/prof_lookup_call_raw.csv 2>/dev/null
tail -2 gpurun_out/ncu_lookup.log
ls -la gpurun_out
