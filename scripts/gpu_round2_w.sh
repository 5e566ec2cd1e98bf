#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_train.py -m gpu -q --timeout 900 > gpurun_out/r2w_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2w_tests.log; tail -3 gpurun_out/r2w_tests.log
NCU="ncu --clock-control none"
python scripts/profile_train.py 64 bf16 > gpurun_out/r2w_plain_train64.log 2>&1 && $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2w_train64_launches.csv python scripts/profile_train.py 64 bf16 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r2w_train64_launches.csv > gpurun_out/r2w_train64_shares.txt 2>&1; grep -E "launches|stem|densify" gpurun_out/r2w_train64_shares.txt | head
