#!/bin/bash
# round 2: warp-per-tile stem v3 (branchy activation): parity, in-situ A/B of warps per CTA, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 > gpurun_out/r2m_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_tests.log; tail -3 gpurun_out/r2m_tests.log
B="python bench.py --no-train --no-sdxl --no-cpu-baseline --no-config5 --no-roofline"
for w in 9 8 9 8; do
TCVN_STEM_WARPS=$w timeout 600 $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('warps $w', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['clocks'])"
done | tee gpurun_out/r2m_ab.txt
TCVN_STEM_WARPS=9 timeout 600 $B --no-overlap-cnns 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('serial cnns', round(d['value']), d['ms_per_step'])" | tee -a gpurun_out/r2m_ab.txt
NCU="ncu --clock-control none"
python scripts/profile_infer.py 256 > gpurun_out/r2m_plain_infer.log 2>&1 && $NCU --set full --import-source on --profile-from-start off -k regex:stem_warp_kernel -s 1 -c 1 -o gpurun_out/r2m_stem -f python scripts/profile_infer.py 256 > gpurun_out/r2m_ncu_stem.log 2>&1
