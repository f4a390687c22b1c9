#!/bin/bash
# GPU box: lookup parity tests once, then the timing of tuning variants of the lookup kernel ("$@" = extra env settings to sweep).
mkdir -p gpurun_out
python -m prior_flow_b200.build > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -x -k "lookup or dccl or corrblock or backward or onthefly or channels_last or e2e" 2>&1 | tail -5
: > gpurun_out/lookup_tune.log
for cfg in "$@"; do
  env $cfg timeout 120 python scripts/lookup_tune.py 2>&1 | tail -1 | tee -a gpurun_out/lookup_tune.log
done
