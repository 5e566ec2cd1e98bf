#!/bin/bash
# round 2, GPU call A: full GPU test suite, determinism, bf16-vs-oracle gate, bench, training launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/r2a_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_tests.log
tail -30 gpurun_out/r2a_tests.log
timeout 300 python scripts/gpu_determinism.py bf16 16 > gpurun_out/r2a_determinism.log 2>&1; tail -5 gpurun_out/r2a_determinism.log
timeout 600 python scripts/gpu_bf16_oracle_gate.py 16 gpurun_out/r2a_gate16.json > gpurun_out/r2a_gate16.log 2>&1; tail -30 gpurun_out/r2a_gate16.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2a_train64_launches.csv python scripts/profile_train.py 64 bf16 > gpurun_out/r2a_ncu_train64.log 2>&1
python scripts/launch_summary.py gpurun_out/r2a_train64_launches.csv > gpurun_out/r2a_train64_shares.txt 2>&1; head -40 gpurun_out/r2a_train64_shares.txt
