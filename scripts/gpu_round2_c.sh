#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sdxl.py -m gpu -q -x --timeout 600 -s > gpurun_out/r2c_sdxl_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_sdxl_tests.log
tail -25 gpurun_out/r2c_sdxl_tests.log
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_umma_train.py -m gpu -q -x --timeout 600 > gpurun_out/r2c_train_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_train_tests.log
tail -5 gpurun_out/r2c_train_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2c_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c_bench.json'))
print('infer', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
for k in ('train','train_large_batch'):
    print(k, d[k]['value'], d[k]['ms_per_step'], d[k]['gpu_launches'])
print('config5', json.dumps(d['config5_max_prongs'])[:600])
print('sdxl', json.dumps(d['sdxl_variant'])[:900])
PY
for ev in 16 64; do
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c_train${ev}_launches.csv python scripts/profile_train.py $ev bf16 > gpurun_out/r2c_ncu_train${ev}.log 2>&1
python scripts/launch_summary.py gpurun_out/r2c_train${ev}_launches.csv > gpurun_out/r2c_train${ev}_shares.txt 2>&1; head -14 gpurun_out/r2c_train${ev}_shares.txt
done
