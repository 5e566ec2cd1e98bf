#!/bin/bash
# conv2: TMA prefetch into L2 N tiles ahead (TCVN_C2_PREFETCH)
T=${1:-r2pf}
mkdir -p gpurun_out
for pf in 0 1 2 3 4 0 2; do
TCVN_C2_PREFETCH=$pf timeout 900 python bench.py --no-cpu-baseline --no-sdxl --no-config5 --no-train 2>gpurun_out/${T}_bench_$pf.err >> gpurun_out/${T}_bench_$pf.json
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench_$pf.json').read().strip().splitlines()[-1])
print('PF=$pf infer', round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']),
      '| conv1 us', round(d['roofline']['us_per_launch'],1), 'conv2 us', round(d['rooflines_other']['conv2']['us_per_launch'],1))
PY
done
