#!/bin/bash
# 2 GPUs of one box: the driver's own launch line for N = 2 (inference replicas + data-parallel training with checksums)
T=${1:-r2n8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline --no-sdxl > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "rc=$?"
tail -3 gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1])
print('n_gpus', d['n_gpus'], 'infer', round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']))
for k in ('train','train_large_batch'):
    t=d[k]; print(k, round(t['value']), round(t['ms_per_step'],2), t.get('data_parallel_check'))
PY
