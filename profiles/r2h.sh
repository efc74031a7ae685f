#!/bin/bash
# round 2, capture h: K1 with two x tiles per CTA (K32_XT = 2) against one (the r2a kernel), bit for bit; tests
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/r2h_pytest.log 2>&1
echo "pytest_exit=$?" | tee -a $OUT/r2h_pytest.log
grep -v "^frame" $OUT/r2h_pytest.log | tail -12
SWEEP_REPS=3 timeout 600 python profiles/sweep_k1.py > $OUT/r2h_sweep_k1.log 2>&1
cat $OUT/r2h_sweep_k1.log
KCMD="python profiles/run_kernels.py 1"
timeout 300 $KCMD > $OUT/r2h_plain_kernels.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'k_anchor_targets' -c 1 \
    -o $OUT/r2h_prof -f $KCMD > $OUT/r2h_ncu_full.log 2>&1
echo "ncufull_exit=$?"
