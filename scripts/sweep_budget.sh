#!/bin/bash
# GPU: inference events/s vs the per-block working-set budget that sizes the CNN chunks (plan.h)
for mb in ${BUDGETS:-48 64 96 128 192 256 384 768}; do
  echo -n "budget ${mb} MB: "
  TCVN_L2_BUDGET_MB=$mb python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train --no-roofline 2>gpurun_out/sweep.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), 'events/s', round(d['ms_per_step'],2), 'ms', d['gpu_launches'], 'launches')" || tail -3 gpurun_out/sweep.err
done
