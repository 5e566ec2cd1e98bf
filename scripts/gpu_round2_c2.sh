#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_train.py -m gpu -q --timeout 900 > gpurun_out/r2c2_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c2_tests.log; tail -3 gpurun_out/r2c2_tests.log
NCU="ncu --clock-control none"
python scripts/profile_train.py 64 bf16 > gpurun_out/r2c2_plain_train64.log 2>&1 && $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2c2_train64_launches.csv python scripts/profile_train.py 64 bf16 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r2c2_train64_launches.csv > gpurun_out/r2c2_train64_shares.txt 2>&1; grep -E "launches|grad_pull|bnact_fwd_vec|bn_param" gpurun_out/r2c2_train64_shares.txt | head
timeout 900 python bench.py --no-sdxl --no-cpu-baseline --no-roofline --no-config5 2>gpurun_out/r2c2_bench.err > gpurun_out/r2c2_bench.json; python -c "
import json; d=json.loads(open('gpurun_out/r2c2_bench.json').read().strip().splitlines()[-1])
print(round(d['value']), d['ms_per_step']); print({k:d['train'].get(k) for k in ('value','ms_per_step','launch_mode')}); print({k:d['train_large_batch'].get(k) for k in ('value','ms_per_step')})"
