#!/bin/bash
# round 2, last evidence run (trimmed to the GPU minutes left): full GPU suite, default bench, reference arm, launch lists, conv1 --set full
T=${1:-r2z3}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/${T}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_tests.log; tail -3 gpurun_out/${T}_tests.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "ref rc=$?"
NCU="ncu --clock-control none"
$NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${T}_infer256_launches.csv python scripts/profile_infer.py 256 > /dev/null 2>&1
for ev in 64 16; do
$NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${T}_train${ev}_launches.csv python scripts/profile_train.py $ev bf16 > /dev/null 2>&1
done
for f in infer256 train16 train64; do python scripts/launch_summary.py gpurun_out/${T}_${f}_launches.csv > gpurun_out/${T}_${f}_shares.txt 2>&1; done
$NCU --set full --import-source on -k regex:umma_gemm_kernel -s 34 -c 3 -o gpurun_out/${T}_conv1 -f python scripts/profile_cnn.py 194 2 --sparse > gpurun_out/${T}_ncu_conv1.log 2>&1
ls gpurun_out/${T}_* | wc -l
