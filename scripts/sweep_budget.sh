#!/bin/bash
# GPU: inference events/s vs the per-block working-set budget that sizes the CNN chunks (plan.h), step replayed as a CUDA graph
mkdir -p gpurun_out
for mb in ${BUDGETS:-96 160 256 512 1024 1536 3072}; do
  echo -n "budget ${mb} MB: "
  TCVN_L2_BUDGET_MB=$mb python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train --no-roofline --no-sdxl --no-config5 2>gpurun_out/sweep.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), 'events/s', round(d['ms_per_step'],2), 'ms', d['gpu_launches'], 'launches', 'e2e', round(d['e2e']['value']))" || tail -3 gpurun_out/sweep.err
done | tee gpurun_out/r2_sweep_budget_graph.txt
