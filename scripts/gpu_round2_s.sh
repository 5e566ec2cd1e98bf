#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_export.py -m gpu -q -x --timeout 600 > gpurun_out/r2s_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s_tests.log; tail -3 gpurun_out/r2s_tests.log
B="python bench.py --no-train --no-sdxl --no-cpu-baseline --no-roofline"
for i in 1 2; do
timeout 600 $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('run $i', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['clocks'], d.get('config5_max_prongs',{}).get('value'))"
done | tee gpurun_out/r2s_ab.txt
NCU="ncu --clock-control none"
python scripts/profile_infer.py 256 > gpurun_out/r2s_plain_infer.log 2>&1 && $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2s_infer256_launches.csv python scripts/profile_infer.py 256 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r2s_infer256_launches.csv > gpurun_out/r2s_infer256_shares.txt 2>&1; head -12 gpurun_out/r2s_infer256_shares.txt
