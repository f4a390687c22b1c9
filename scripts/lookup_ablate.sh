#!/bin/bash
mkdir -p gpurun_out
python -m prior_flow_b200.build > gpurun_out/build.log 2>&1
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"lookup|rotate" -s 6 -c 2 --csv python scripts/lookup_tune.py --reps 1 2>/dev/null | grep -E "lookup|rotate" | awk -F'","' '{print $5, $(NF-2), $NF}' 
done
