#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --no-train --no-sdxl --no-cpu-baseline --no-config5 --no-roofline"
for v in 0 1 0 1; do
if [ $v = 1 ]; then export TCVN_STEM_SKIP=1; else unset TCVN_STEM_SKIP; fi
timeout 600 $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('skip $v', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['clocks'])"
done | tee gpurun_out/r2n_ab.txt
unset TCVN_STEM_SKIP
timeout 600 $B --steps 60 --warmup 10 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('60 steps', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['clocks'])" | tee -a gpurun_out/r2n_ab.txt
