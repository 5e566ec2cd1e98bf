#!/bin/bash
T=${1:-r2acc4}
mkdir -p gpurun_out
for k in 64 128; do
TCVN_C1_TRACE=gpurun_out/${T}_k$k.txt TCVN_C1_TRACE_K=$k python scripts/profile_cnn.py 194 2 --sparse > gpurun_out/${T}_k$k.log 2>&1
done
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
