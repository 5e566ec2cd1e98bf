#!/bin/bash
T=${1:-r2trace}
mkdir -p gpurun_out
for k in 64 128; do
TCVN_C1_TRACE=gpurun_out/${T}_k$k.txt TCVN_C1_TRACE_K=$k python scripts/profile_cnn.py 194 2 --sparse > gpurun_out/${T}_k$k.log 2>&1
done
ls -la gpurun_out/${T}_*
