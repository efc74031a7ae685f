#!/bin/bash
# round 2, capture e: K3 stream with 512-score warp tiles; NMS sweep between rounds; mailbox connect
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r2e_pytest.log 2>&1
echo "pytest_exit=$?" | tee -a $OUT/r2e_pytest.log
tail -30 $OUT/r2e_pytest.log
timeout 300 python profiles/nms_phases.py > $OUT/r2e_nms_phases.log 2>&1
cat $OUT/r2e_nms_phases.log
timeout 300 python profiles/time_inference.py > $OUT/r2e_time_inference.log 2>&1
cat $OUT/r2e_time_inference.log
KCMD="python profiles/run_kernels.py 1"
timeout 300 $KCMD > $OUT/r2e_plain_kernels.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'k_threshold_keys|k_segment_nms|k_merge_topk' -c 3 \
    -o $OUT/r2e_prof -f $KCMD > $OUT/r2e_ncu_full.log 2>&1
echo "ncufull_exit=$?"
tail -3 $OUT/r2e_ncu_full.log
