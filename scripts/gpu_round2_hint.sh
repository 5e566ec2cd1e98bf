#!/bin/bash
# L2 eviction hints on the conv1 kernel's TMA loads: DRAM bytes (ncu) and timings
T=${1:-r2hint}
mkdir -p gpurun_out
for h in 0 1 3; do
TCVN_TMA_HINT=$h ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:umma_gemm_kernel -s 34 -c 7 --csv --log-file gpurun_out/${T}_ncu_bytes_$h.csv python scripts/profile_cnn.py 194 2 --sparse > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/${T}_ncu_bytes_$h.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
col={h:i for i,h in enumerate(rows[hi])}
cur={}
for r in rows[hi+1:]:
    if len(r)<len(col): continue
    cur.setdefault(r[col['ID']],{})[r[col['Metric Name']]]=r[col['Metric Value']]
for k,v in cur.items(): print('hint=$h',k,{a.split('__')[-1][:20]:b for a,b in v.items()})
PY
done
for h in 0 1 3 0 1 3; do
TCVN_TMA_HINT=$h timeout 900 python bench.py --no-cpu-baseline --no-train --no-sdxl --no-config5 2>gpurun_out/${T}_bench_$h.err >> gpurun_out/${T}_bench_$h.json
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench_$h.json').read().strip().splitlines()[-1])
print('HINT=$h infer', round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']),
      '| conv1 us', round(d['roofline']['us_per_launch'],1), 'conv2 us', round(d['rooflines_other']['conv2']['us_per_launch'],1))
PY
done
