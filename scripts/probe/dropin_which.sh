# the full default bench run: drop-in stage split (build_pyramid was 5.6 ms while auto mode asked the driver for free memory per forward)
python bench.py 2>/dev/null | tail -1 > gpurun_out/r03z_bench_line.json
python -c "
import json;d=json.load(open('gpurun_out/r03z_bench_line.json'));e=d['dropin']['eager'];print(d['value'], d['e2e']['value'], e['value'], e['hot_path_ms'], e['stages_ms']['pyramid (DCCL.build_pyramid)'], e['stages_ms']['lookup (DCCL.__call__)'], d['dropin']['cuda_graph']['value'])"
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r03z_bench_reference_line.json; cut -c1-200 gpurun_out/r03z_bench_reference_line.json
