#!/bin/bash
# round 2, evidence run: full GPU suite, default bench (what the driver runs), launch lists, --set full captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2h_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_tests.log; tail -4 gpurun_out/r2h_tests.log
timeout 900 python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2h_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2h_bench_ref.json 2> gpurun_out/r2h_bench_ref.err; echo "ref rc=$?"
NCU="ncu --clock-control none"
python scripts/profile_infer.py 256 > gpurun_out/r2h_plain_infer.log 2>&1 && $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2h_infer256_launches.csv python scripts/profile_infer.py 256 > /dev/null 2>&1
for ev in 16 64; do
python scripts/profile_train.py $ev bf16 > gpurun_out/r2h_plain_train$ev.log 2>&1 && $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2h_train${ev}_launches.csv python scripts/profile_train.py $ev bf16 > /dev/null 2>&1
done
python scripts/profile_sdxl.py 6 > gpurun_out/r2h_plain_sdxl.log 2>&1 && $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2h_sdxl6_launches.csv python scripts/profile_sdxl.py 6 > /dev/null 2>&1
# --set full captures
python scripts/profile_cnn.py 194 2 --sparse > gpurun_out/r2h_plain_cnn.log 2>&1 && $NCU --set full --import-source on -k regex:umma_gemm_kernel -s 34 -c 3 -o gpurun_out/r2h_conv1 -f python scripts/profile_cnn.py 194 2 --sparse > gpurun_out/r2h_ncu_conv1.log 2>&1
$NCU --set full --import-source on --profile-from-start off -k regex:umma_conv2_dgrad -s 57 -c 2 -o gpurun_out/r2h_dgrad -f python scripts/profile_train.py 64 bf16 > gpurun_out/r2h_ncu_dgrad.log 2>&1
$NCU --set full --import-source on --profile-from-start off -k regex:grad_pull_kernel -s 57 -c 2 -o gpurun_out/r2h_pull -f python scripts/profile_train.py 64 bf16 > gpurun_out/r2h_ncu_pull.log 2>&1
$NCU --set full --import-source on --profile-from-start off -k regex:umma_gemm_kernel -s 3 -c 2 -o gpurun_out/r2h_sdxlconv -f python scripts/profile_sdxl.py 6 > gpurun_out/r2h_ncu_sdxlconv.log 2>&1
$NCU --set full --import-source on --profile-from-start off -k regex:stem_train_kernel -c 2 -o gpurun_out/r2h_stemtrain -f python scripts/profile_train.py 64 bf16 > gpurun_out/r2h_ncu_stemtrain.log 2>&1
ls -la gpurun_out/r2h_* | head -40
