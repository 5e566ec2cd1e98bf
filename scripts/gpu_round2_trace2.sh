#!/bin/bash
T=${1:-r2trace2}
mkdir -p gpurun_out
TCVN_C2_TRACE=gpurun_out/${T}_c2.txt python scripts/profile_cnn.py 194 2 --sparse > gpurun_out/${T}_c2.log 2>&1
ls -la gpurun_out/${T}_*
