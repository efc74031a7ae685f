#!/bin/bash
# GPU-box capture recipe (run under gpurun from the repo root):
#   gpurun --timeout 2400 -- 'bash profiles/capture.sh r2z'
# 1. parity tests  2. the default bench line  3. ncu launch lists of the bench command (value region; whole command)
# 4. one `ncu --set full` capture of the hot kernels (profiles/run_kernels.py, no CUDA graphs).
# Every ncu run follows a plain run of the same command that exited 0.
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1
echo "pytest_exit=$?" | tee -a $OUT/pytest_$TAG.log
grep -v "^frame" $OUT/pytest_$TAG.log | tail -4
python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench_exit=$?"
head -c 600 $OUT/bench_$TAG.json; echo
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err
echo "bench_ref_exit=$?"
# launch list of the `value` region only (K1 + K2 per step), then of the whole default bench command
VCMD="python bench.py --steps 10 --warmup 3 --no-cpu --no-extras --no-e2e"
$VCMD > $OUT/plain_value_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv \
    --log-file $OUT/launches_value_$TAG.csv $VCMD > $OUT/ncu_launch_value_$TAG.log 2>&1
echo "launchlist_value_exit=$?"
BCMD="python bench.py --steps 3 --warmup 3 --no-cpu --config 2 --no-extras"
$BCMD > $OUT/plain_infer_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file $OUT/launches_infer_$TAG.csv $BCMD > $OUT/ncu_launch_infer_$TAG.log 2>&1
echo "launchlist_infer_exit=$?"
KCMD="python profiles/run_kernels.py 2"
$KCMD > $OUT/plain_kernels_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on \
    -k regex:'k_anchor_targets|k_loss|k_threshold_keys|k_segment_nms|k_merge_topk' -c 10 \
    -o $OUT/prof_$TAG -f $KCMD > $OUT/ncu_full_$TAG.log 2>&1
echo "ncufull_exit=$?"
tail -3 $OUT/ncu_full_$TAG.log
