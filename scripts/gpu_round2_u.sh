#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_train.py -m gpu -q --timeout 900 > gpurun_out/r2u_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2u_tests.log; tail -5 gpurun_out/r2u_tests.log
python scripts/gpu_bf16_oracle_gate.py 16 gpurun_out/r2u_gate16.json > gpurun_out/r2u_gate16.log 2>&1; cat gpurun_out/r2u_gate16.json | head -40
