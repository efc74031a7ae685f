#!/bin/bash
# One `ncu --set full` capture of the hot kernels, after a plain run of the same command:  bash profiles/ncu_kernels.sh <tag> [regex]
set -u
TAG=${1:-r2}
RE=${2:-'k_anchor_targets|k_loss|k_threshold_keys|k_segment_nms|k_merge_topk'}
OUT=gpurun_out
mkdir -p $OUT
KCMD="python profiles/run_kernels.py 2"
$KCMD > $OUT/plain_kernels_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$RE" -c 10 \
    -o $OUT/prof_$TAG -f $KCMD > $OUT/ncu_full_$TAG.log 2>&1
echo "ncufull_exit=$?"
tail -3 $OUT/ncu_full_$TAG.log
