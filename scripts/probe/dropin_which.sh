# does the full default bench run reproduce the slow build_pyramid stage of the drop-in leg?
for flags in "" "--steps 5"; do
python bench.py $flags 2>/dev/null | tail -1 > gpurun_out/which.json
python -c "
import json;d=json.load(open('gpurun_out/which.json'));e=d['dropin']['eager'];print('flags [$flags]', d['value'], e['value'], e['hot_path_ms'], e['stages_ms']['pyramid (DCCL.build_pyramid)'], e['stages_ms']['lookup (DCCL.__call__)'])"
done
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
