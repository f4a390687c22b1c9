#!/bin/bash
# tensor-core on-the-fly lookup: parity tests, tile statistics, kernel times, per-kernel launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_onthefly_tc.py -x -q -s > gpurun_out/otf_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/otf_tests.log
grep -v "^$" gpurun_out/otf_tests.log | tail -30
timeout 300 python scripts/probe/otf_tiles.py > gpurun_out/otf_tiles.txt 2>&1; cat gpurun_out/otf_tiles.txt
timeout 600 python scripts/kbench.py --only onthefly --out gpurun_out/otf_kbench.json > gpurun_out/otf_kbench.log 2>&1
tail -8 gpurun_out/otf_kbench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"otf_|rotate" --csv --log-file gpurun_out/otf_launches.csv python scripts/probe/otf_tiles.py > /dev/null 2>&1
