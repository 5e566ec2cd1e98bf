#!/bin/bash
# GEMM kernel without the statistics accumulators in the non-shift-MMA instantiations (register spills): training tests + per-kernel totals
T=${1:-r2spill}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_umma_train.py tests/test_gpu_sdxl.py -m gpu -q --timeout 600 -x > gpurun_out/${T}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_tests.log; tail -3 gpurun_out/${T}_tests.log
ncu --clock-control none --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${T}_train64.csv python scripts/profile_train.py 64 bf16 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/${T}_train64.csv | grep -E "launches|umma_"
