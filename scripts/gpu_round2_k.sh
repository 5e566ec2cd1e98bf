#!/bin/bash
# round 2: warp-per-tile stem - parity, bench, launch list, ncu capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_export.py -m gpu -q -x --timeout 600 > gpurun_out/r2k_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_tests.log; tail -6 gpurun_out/r2k_tests.log
timeout 600 python bench.py --no-train --no-sdxl --no-cpu-baseline > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2k_bench.err
NCU="ncu --clock-control none"
python scripts/profile_infer.py 256 > gpurun_out/r2k_plain_infer.log 2>&1 && $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2k_infer256_launches.csv python scripts/profile_infer.py 256 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r2k_infer256_launches.csv > gpurun_out/r2k_infer256_shares.txt 2>&1; head -20 gpurun_out/r2k_infer256_shares.txt
$NCU --set full --import-source on --profile-from-start off -k regex:stem_warp_kernel -s 1 -c 1 -o gpurun_out/r2k_stem -f python scripts/profile_infer.py 256 > gpurun_out/r2k_ncu_stem.log 2>&1
ls -la gpurun_out/r2k_*
