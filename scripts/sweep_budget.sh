#!/bin/bash
# GPU: events/s vs the per-block L2 budget that sizes the CNN chunks
for mb in ${BUDGETS:-48 64 80 100 128 160}; do
  echo -n "budget ${mb} MB: "
  TCVN_L2_BUDGET_MB=$mb python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), 'events/s', round(d['ms_per_step'],2), 'ms')"
done
