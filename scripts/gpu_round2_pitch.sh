#!/bin/bash
# row pitch of the eval block buffers padded to whole 128-byte lines (TCVN_PAD_PITCH=1 default / 0): tests, A/B bench, DRAM bytes
T=${1:-r2pitch}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/${T}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_tests.log; tail -3 gpurun_out/${T}_tests.log
for pad in 1 0 1 0; do
TCVN_PAD_PITCH=$pad timeout 900 python bench.py --no-cpu-baseline --no-train --no-sdxl 2>gpurun_out/${T}_bench_$pad.err >> gpurun_out/${T}_bench_$pad.json
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench_$pad.json').read().strip().splitlines()[-1])
print('PAD=$pad infer', round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']),
      '| conv1 us', round(d['roofline']['us_per_launch'],1), 'conv2 us', round(d['rooflines_other']['conv2']['us_per_launch'],1),
      '| cfg5', round(d['config5_max_prongs']['inference']['ms_per_step'],2), '| 1ev', round(d['single_event_latency']['graph_us']))
PY
done
for pad in 1 0; do
TCVN_PAD_PITCH=$pad ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:umma_gemm_kernel -s 34 -c 7 --csv --log-file gpurun_out/${T}_ncu_bytes_$pad.csv python scripts/profile_cnn.py 194 2 --sparse > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/${T}_ncu_bytes_$pad.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
col={h:i for i,h in enumerate(rows[hi])}
cur={}
for r in rows[hi+1:]:
    if len(r)<len(col): continue
    cur.setdefault(r[col['ID']],{})[r[col['Metric Name']]]=r[col['Metric Value']]
for k,v in cur.items(): print('pad=$pad',k,{a.split('__')[-1][:20]:b for a,b in v.items()})
PY
done
