#!/bin/bash
# per-map L2 promotion rule vs forced settings, and resident weights on/off: per-kernel totals of one training step (64 events)
# and one inference step (256 events), ncu time-only launch lists
T=${1:-r2promo3}
mkdir -p gpurun_out
NCU="ncu --clock-control none --profile-from-start off --metrics gpu__time_duration.sum --csv"
for cfg in "auto 1" "0 1" "256 1" "auto 0"; do
set -- $cfg
if [ "$1" = "auto" ]; then unset TCVN_TMAP_PROMO; else export TCVN_TMAP_PROMO=$1; fi
export TCVN_C1_WRES=$2
$NCU --log-file gpurun_out/${T}_train64_$1_$2.csv python scripts/profile_train.py 64 bf16 > /dev/null 2>&1
$NCU --log-file gpurun_out/${T}_infer_$1_$2.csv python scripts/profile_infer.py 256 > /dev/null 2>&1
for w in train64 infer; do
echo "== $w promo=$1 wres=$2"; python scripts/launch_summary.py gpurun_out/${T}_${w}_$1_$2.csv | grep -E "launches|umma_|act_pool2" 
done
done
