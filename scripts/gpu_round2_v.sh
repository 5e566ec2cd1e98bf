#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --clock-control none"
python scripts/profile_train.py 64 bf16 > gpurun_out/r2v_plain_train64.log 2>&1 && $NCU --set full --import-source on --profile-from-start off -k regex:"stem_train|stem_pool16" -c 8 -o gpurun_out/r2v_stemtrain -f python scripts/profile_train.py 64 bf16 > gpurun_out/r2v_ncu.log 2>&1
tail -2 gpurun_out/r2v_ncu.log
