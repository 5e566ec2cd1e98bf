#!/bin/bash
# padded row pitch x L2 promotion of the tensor maps: DRAM bytes of the conv1 launches (ncu) and step timings
T=${1:-r2pitch2}
mkdir -p gpurun_out
for cfg in "1 0" "1 128"; do
set -- $cfg
TCVN_PAD_PITCH=$1 TCVN_TMAP_PROMO=$2 ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:umma_gemm_kernel -s 34 -c 7 --csv --log-file gpurun_out/${T}_ncu_bytes_$1_$2.csv python scripts/profile_cnn.py 194 2 --sparse > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/${T}_ncu_bytes_$1_$2.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
col={h:i for i,h in enumerate(rows[hi])}
cur={}
for r in rows[hi+1:]:
    if len(r)<len(col): continue
    cur.setdefault(r[col['ID']],{})[r[col['Metric Name']]]=r[col['Metric Value']]
for k,v in cur.items(): print('pad=$1 promo=$2',k,{a.split('__')[-1][:20]:b for a,b in v.items()})
PY
done
for cfg in "0 256" "1 0" "1 128" "0 256" "1 0" "1 128"; do
set -- $cfg
TCVN_PAD_PITCH=$1 TCVN_TMAP_PROMO=$2 timeout 900 python bench.py --no-cpu-baseline --no-sdxl --no-config5 --no-train 2>gpurun_out/${T}_bench_$1_$2.err >> gpurun_out/${T}_bench_$1_$2.json
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench_$1_$2.json').read().strip().splitlines()[-1])
print('PAD=$1 PROMO=$2 infer', round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']),
      '| conv1 us', round(d['roofline']['us_per_launch'],1), 'conv2 us', round(d['rooflines_other']['conv2']['us_per_launch'],1))
PY
done
