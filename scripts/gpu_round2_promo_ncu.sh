#!/bin/bash
# DRAM bytes of the first six conv1 launches of dense block 1 under the three L2-promotion settings of the tensor maps
T=${1:-r2promo}
mkdir -p gpurun_out
for promo in 256 0; do
TCVN_TMAP_PROMO=$promo ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum -k regex:umma_gemm_kernel -s 34 -c 7 --csv --log-file gpurun_out/${T}_ncu_bytes_$promo.csv python scripts/profile_cnn.py 194 2 --sparse > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/${T}_ncu_bytes_$promo.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
col={h:i for i,h in enumerate(rows[hi])}
cur={}
for r in rows[hi+1:]:
    if len(r)<len(col): continue
    cur.setdefault(r[col['ID']],{})[r[col['Metric Name']]]=(r[col['Metric Value']],r[col['Metric Unit']])
for k,v in cur.items(): print('promo=$promo',k,{a.split('__')[-1][:40]:b for a,b in v.items()})
PY
done
