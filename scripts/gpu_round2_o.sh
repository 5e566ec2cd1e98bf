#!/bin/bash
# round 2: full GPU suite with the new stem + ncu of the fused SDXL 2-D conv kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2o_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_tests.log; tail -4 gpurun_out/r2o_tests.log
NCU="ncu --clock-control none"
python scripts/profile_sdxl.py 48 > gpurun_out/r2o_plain_sdxl.log 2>&1 && $NCU --set full --import-source on --profile-from-start off -k regex:umma_conv2d_c64 -s 4 -c 2 -o gpurun_out/r2o_conv2d -f python scripts/profile_sdxl.py 48 > gpurun_out/r2o_ncu_conv2d.log 2>&1
tail -3 gpurun_out/r2o_ncu_conv2d.log
