#!/bin/bash
# GPU box: bench (our arm) + kbench — a mid-round checkpoint.
mkdir -p gpurun_out
python -m prior_flow_b200.build > gpurun_out/build.log 2>&1
timeout 900 python bench.py --skip-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -1 gpurun_out/bench.log | cut -c1-2400; tail -3 gpurun_out/bench.err
timeout 300 python scripts/kbench.py --iters 20 --skip-torch 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: r=json.loads(l); print(r['kernel'], r['ms_median'], r.get('GBps'), r.get('frac_hbm'))
    except Exception: pass"
