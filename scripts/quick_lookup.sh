#!/bin/bash
# GPU box: correctness of the lookup tests + duration / instruction count of the lookup kernels (ncu, light metric set)
mkdir -p gpurun_out
python -m prior_flow_b200.build > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -k "lookup or dccl or corrblock or backward" 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum \
  --clock-control none -k regex:"lookup|rotate" -c 6 --csv --log-file gpurun_out/quick_lookup.csv python scripts/kbench.py --iters 1 --skip-torch > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/quick_lookup.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[hi]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
cur={}
for r in rows[hi+1:]:
    cur.setdefault((r[ii],r[ki][:34]),{})[r[mi]]=r[vi]
for k,v in cur.items(): print(k, {a.split('.')[0].replace('smsp__','').replace('sm__','')[:22]:b for a,b in v.items()})
PY
timeout 300 python scripts/kbench.py --iters 10 --skip-torch 2>&1 | grep -E "lookup_dual|lookup_single"
