#!/bin/bash
T=${1:-r2promo2}
mkdir -p gpurun_out
for promo in 0 256 0 256; do
TCVN_TMAP_PROMO=$promo timeout 900 python bench.py --no-cpu-baseline --sdxl-events 64 2>gpurun_out/${T}_bench_$promo.err >> gpurun_out/${T}_bench_$promo.json
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench_$promo.json').read().strip().splitlines()[-1])
print('PROMO=$promo infer', round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']),
      '| conv1 us', round(d['roofline']['us_per_launch'],1), 'conv2 us', round(d['rooflines_other']['conv2']['us_per_launch'],1),
      '| train16', round(d['train']['ms_per_step'],2), 'train64', round(d['train_large_batch']['ms_per_step'],2),
      '| cfg5', round(d['config5_max_prongs']['inference']['ms_per_step'],2), round(d['config5_max_prongs']['training']['ms_per_step'],2),
      '| sdxl', round(d['sdxl_variant']['value']))
PY
done
