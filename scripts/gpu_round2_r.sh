#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sdxl.py -m gpu -q -x --timeout 600 > gpurun_out/r2r_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2r_tests.log; tail -3 gpurun_out/r2r_tests.log
timeout 900 python bench.py --no-train --no-cpu-baseline --no-config5 --no-roofline 2>gpurun_out/r2r_bench.err > gpurun_out/r2r_bench.json; python -c "import json; d=json.loads(open('gpurun_out/r2r_bench.json').read().strip().splitlines()[-1]); print(d.get('sdxl_variant'))"
NCU="ncu --clock-control none"
python scripts/profile_sdxl.py 48 > gpurun_out/r2r_plain_sdxl.log 2>&1 && $NCU --set full --import-source on --profile-from-start off -k regex:umma_conv2d_c64 -s 4 -c 2 -o gpurun_out/r2r_conv2d -f python scripts/profile_sdxl.py 48 > gpurun_out/r2r_ncu_conv2d.log 2>&1
tail -2 gpurun_out/r2r_ncu_conv2d.log
