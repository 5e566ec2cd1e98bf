#!/bin/bash
# stream-overlap policy and chunk budget after the conv1 change
T=${1:-r2misc}
mkdir -p gpurun_out
run() { # label, env..., args
  label=$1; shift
  env "$@" timeout 900 python bench.py --no-cpu-baseline --no-sdxl --no-config5 --no-train --no-roofline $EXTRA 2>gpurun_out/${T}_bench.err | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$label', round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), d['gpu_launches'])"
}
for i in 1 2; do
run base A=1
EXTRA=--no-overlap-cnns run no-overlap A=1
run budget3072 TCVN_L2_BUDGET_MB=3072
run budget6144 TCVN_L2_BUDGET_MB=6144
run budget768 TCVN_L2_BUDGET_MB=768
done
