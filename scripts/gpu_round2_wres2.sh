#!/bin/bash
# conv1 with resident weights + deeper A pipeline + four accumulators, padded pitch, no L2 promotion: tests, bench x2, DRAM bytes
T=${1:-r2wres2}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/${T}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_tests.log; tail -3 gpurun_out/${T}_tests.log
for i in 1 2; do
timeout 900 python bench.py --no-cpu-baseline 2>gpurun_out/${T}_bench.err >> gpurun_out/${T}_bench.json
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1])
print('infer', round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']),
      '| conv1 us', round(d['roofline']['us_per_launch'],1), 'conv2 us', round(d['rooflines_other']['conv2']['us_per_launch'],1),
      '| train16', round(d['train']['ms_per_step'],2), 'train64', round(d['train_large_batch']['ms_per_step'],2),
      '| cfg5', round(d['config5_max_prongs']['inference']['ms_per_step'],2), round(d['config5_max_prongs']['training']['ms_per_step'],2),
      '| sdxl', round(d['sdxl_variant']['value']))
PY
done
ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:umma_gemm_kernel -s 34 -c 7 --csv --log-file gpurun_out/${T}_ncu_bytes.csv python scripts/profile_cnn.py 194 2 --sparse > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/${T}_ncu_bytes.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
col={h:i for i,h in enumerate(rows[hi])}
cur={}
for r in rows[hi+1:]:
    if len(r)<len(col): continue
    cur.setdefault(r[col['ID']],{})[r[col['Metric Name']]]=r[col['Metric Value']]
for k,v in cur.items(): print(k,{a.split('__')[-1][:20]:b for a,b in v.items()})
PY
