#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_umma_train.py -m gpu -q -x --timeout 600 -s > gpurun_out/r2e_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_tests.log
grep -E "passed|failed|rc=|tcvn_bf16|Error|error" gpurun_out/r2e_tests.log | tail -8
timeout 300 python scripts/gpu_train_host_bound.py 16 64 2>&1 | tail -3
for ev in 16 64; do
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2e_train${ev}_launches.csv python scripts/profile_train.py $ev bf16 > gpurun_out/r2e_ncu_train${ev}.log 2>&1
python scripts/launch_summary.py gpurun_out/r2e_train${ev}_launches.csv > gpurun_out/r2e_train${ev}_shares.txt 2>&1; head -24 gpurun_out/r2e_train${ev}_shares.txt
done
