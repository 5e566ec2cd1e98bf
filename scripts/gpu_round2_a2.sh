#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sdxl.py -m gpu -q -x --timeout 600 > gpurun_out/r2a2_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a2_tests.log; tail -8 gpurun_out/r2a2_tests.log
timeout 900 python bench.py --no-train --no-cpu-baseline --no-config5 --no-roofline 2>gpurun_out/r2a2_bench.err > gpurun_out/r2a2_bench.json; python -c "import json; d=json.loads(open('gpurun_out/r2a2_bench.json').read().strip().splitlines()[-1]); s=d.get('sdxl_variant'); print(s['value'], s['ms_per_step'])"
