#!/bin/bash
# GPU: events/s vs the image count of a block-0 (stem + dense1) chunk; later blocks keep the default budget
for c in ${CHUNKS:-20 24 32 40 48 64 96 128}; do
  echo -n "chunk0 ${c}: "
  TCVN_CHUNK0=$c python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), 'events/s', round(d['ms_per_step'],2), 'ms')"
done
