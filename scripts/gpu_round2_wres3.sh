#!/bin/bash
# epilogue hoists in the conv1 kernel; A/B of the resident-weight mode on the training legs
T=${1:-r2wres3}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_train.py -m gpu -q --timeout 900 -x > gpurun_out/${T}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_tests.log; tail -3 gpurun_out/${T}_tests.log
for w in 1 0 1 0; do
TCVN_C1_WRES=$w timeout 900 python bench.py --no-cpu-baseline --no-sdxl --no-config5 2>gpurun_out/${T}_bench_$w.err >> gpurun_out/${T}_bench_$w.json
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench_$w.json').read().strip().splitlines()[-1])
print('WRES=$w infer', round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']),
      '| conv1 us', round(d['roofline']['us_per_launch'],1), 'conv2 us', round(d['rooflines_other']['conv2']['us_per_launch'],1),
      '| train16', round(d['train']['ms_per_step'],2), 'train64', round(d['train_large_batch']['ms_per_step'],2))
PY
done
ncu --clock-control none --metrics gpu__time_duration.sum -k regex:umma_gemm_kernel -s 34 -c 3 --csv --log-file gpurun_out/${T}_ncu_t.csv python scripts/profile_cnn.py 194 2 --sparse > /dev/null 2>&1
grep -o '"gpu__time_duration.sum","[a-z]*","[0-9.,]*"' gpurun_out/${T}_ncu_t.csv
