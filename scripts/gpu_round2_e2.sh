#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r2e2_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e2_tests.log; tail -3 gpurun_out/r2e2_tests.log
NCU="ncu --clock-control none"
for ev in 16 64; do
python scripts/profile_train.py $ev bf16 > gpurun_out/r2e2_plain_train$ev.log 2>&1 && $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2e2_train${ev}_launches.csv python scripts/profile_train.py $ev bf16 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r2e2_train${ev}_launches.csv > gpurun_out/r2e2_train${ev}_shares.txt 2>&1; grep -E "launches|simt_gemm|wgrad_kernel<float" gpurun_out/r2e2_train${ev}_shares.txt
done
timeout 900 python bench.py --no-sdxl --no-cpu-baseline --no-roofline --no-config5 2>gpurun_out/r2e2_bench.err > gpurun_out/r2e2_bench.json; python -c "
import json; d=json.loads(open('gpurun_out/r2e2_bench.json').read().strip().splitlines()[-1])
print(round(d['value']), d['ms_per_step']); print({k:d['train'].get(k) for k in ('value','ms_per_step')}); print({k:d['train_large_batch'].get(k) for k in ('value','ms_per_step')})"
