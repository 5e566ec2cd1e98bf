#!/bin/bash
# per-launch DRAM bytes / duration / issue utilisation of the elementwise backward kernels (all sizes)
mkdir -p gpurun_out
python scripts/profile_train.py 64 bf16 > gpurun_out/r2d2_plain.log 2>&1 || exit 1
ncu --clock-control none --profile-from-start off -k regex:"bnact_bwd_apply16|colsum_vec_kernel|bn1_bwd_reduce_vec|grad_pull_kernel|bnact_fwd_vec|pool2_bwd|act_pool2" \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active \
  --csv --log-file gpurun_out/r2d2_elementwise.csv python scripts/profile_train.py 64 bf16 > /dev/null 2>&1
wc -l gpurun_out/r2d2_elementwise.csv
