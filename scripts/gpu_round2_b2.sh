#!/bin/bash
# round 2: are the HBM-bound backward passes of dense block 1 memory- or issue-bound?
mkdir -p gpurun_out
NCU="ncu --clock-control none --set full --import-source on --profile-from-start off"
python scripts/profile_train.py 64 bf16 > gpurun_out/r2b2_plain.log 2>&1 || exit 1
$NCU -k regex:colsum_vec_kernel -s 44 -c 4 -o gpurun_out/r2b2_colsum -f python scripts/profile_train.py 64 bf16 > gpurun_out/r2b2_ncu1.log 2>&1
$NCU -k regex:bn1_bwd_reduce_vec -s 56 -c 3 -o gpurun_out/r2b2_bn1 -f python scripts/profile_train.py 64 bf16 > gpurun_out/r2b2_ncu2.log 2>&1
$NCU -k regex:bnact_bwd_apply_vec -s 66 -c 5 -o gpurun_out/r2b2_apply -f python scripts/profile_train.py 64 bf16 > gpurun_out/r2b2_ncu3.log 2>&1
ls -la gpurun_out/r2b2_*
