#!/bin/bash
# round 2, final evidence run: full GPU suite, default bench (what the driver runs), reference arm, launch lists, --set full captures
T=${1:-r2z}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/${T}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_tests.log; tail -4 gpurun_out/${T}_tests.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "ref rc=$?"
NCU="ncu --clock-control none"
python scripts/profile_infer.py 256 > gpurun_out/${T}_plain_infer.log 2>&1 && $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${T}_infer256_launches.csv python scripts/profile_infer.py 256 > /dev/null 2>&1
for ev in 16 64; do
python scripts/profile_train.py $ev bf16 > gpurun_out/${T}_plain_train$ev.log 2>&1 && $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${T}_train${ev}_launches.csv python scripts/profile_train.py $ev bf16 > /dev/null 2>&1
done
python scripts/profile_sdxl.py 48 > gpurun_out/${T}_plain_sdxl.log 2>&1 && $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${T}_sdxl48_launches.csv python scripts/profile_sdxl.py 48 > /dev/null 2>&1
for f in infer256 train16 train64 sdxl48; do python scripts/launch_summary.py gpurun_out/${T}_${f}_launches.csv > gpurun_out/${T}_${f}_shares.txt 2>&1; done
# --set full captures
python scripts/profile_cnn.py 194 2 --sparse > gpurun_out/${T}_plain_cnn.log 2>&1 && $NCU --set full --import-source on -k regex:umma_gemm_kernel -s 34 -c 3 -o gpurun_out/${T}_conv1 -f python scripts/profile_cnn.py 194 2 --sparse > gpurun_out/${T}_ncu_conv1.log 2>&1
$NCU --set full --import-source on --profile-from-start off -k regex:"stem_warp_kernel|seq_forward" -c 3 -o gpurun_out/${T}_stem_seq -f python scripts/profile_infer.py 256 > gpurun_out/${T}_ncu_stem_seq.log 2>&1
$NCU --set full --import-source on --profile-from-start off -k regex:"grad_pull_kernel|stem_train" -s 2 -c 5 -o gpurun_out/${T}_train -f python scripts/profile_train.py 64 bf16 > gpurun_out/${T}_ncu_train.log 2>&1
$NCU --set full --import-source on --profile-from-start off -k regex:umma_conv2d_c64 -s 4 -c 2 -o gpurun_out/${T}_conv2d -f python scripts/profile_sdxl.py 48 > gpurun_out/${T}_ncu_conv2d.log 2>&1
ls gpurun_out/${T}_* | wc -l
