#!/bin/bash
# A/B of programmatic dependent launch (TCVN_PDL=1 default / 0): full GPU suite with it on, then the bench legs both ways
T=${1:-r2pdl}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/${T}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_tests.log; tail -3 gpurun_out/${T}_tests.log
for pdl in 1 0 1 0; do
TCVN_PDL=$pdl timeout 900 python bench.py --no-cpu-baseline --no-roofline --sdxl-events 64 2>gpurun_out/${T}_bench_pdl$pdl.err >> gpurun_out/${T}_bench_pdl$pdl.json
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench_pdl$pdl.json').read().strip().splitlines()[-1])
print('PDL=$pdl infer', round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']),
      '| train16', round(d['train']['ms_per_step'],2), 'train64', round(d['train_large_batch']['ms_per_step'],2),
      '| cfg5', round(d['config5_max_prongs']['inference']['ms_per_step'],2), round(d['config5_max_prongs']['training']['ms_per_step'],2),
      '| sdxl', round(d['sdxl_variant']['value']), '| 1ev', round(d['single_event_latency']['graph_us']))
PY
done
