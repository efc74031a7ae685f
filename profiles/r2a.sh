#!/bin/bash
# round 2, capture a: new K1 (y-target table) -- parity, A/B sweep against the round-1 builds, NMS phase baseline,
# source-level ncu of K1.   gpurun --timeout 1200 -- 'bash profiles/r2a.sh'
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/r2a_smi.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/r2a_pytest.log 2>&1
echo "pytest_exit=$?" | tee -a $OUT/r2a_pytest.log
tail -5 $OUT/r2a_pytest.log
SWEEP_REPS=2 timeout 600 python profiles/sweep_k1.py > $OUT/r2a_sweep_k1.log 2>&1
cat $OUT/r2a_sweep_k1.log
timeout 300 python profiles/nms_phases.py > $OUT/r2a_nms_phases.log 2>&1
cat $OUT/r2a_nms_phases.log
timeout 600 python bench.py --steps 20 --no-cpu > $OUT/r2a_bench.json 2> $OUT/r2a_bench.err
echo "bench_exit=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2a_bench.json"))
    print({k: d[k] for k in ("value", "ms_per_step")}, d["roofline"]["us_per_launch"], d["roofline_k2"]["us_per_launch"], d["inference"]["reference_semantics"])
except Exception as e:
    print("bench parse failed", e)
PY
KCMD="python profiles/run_kernels.py 1"
timeout 300 $KCMD > $OUT/r2a_plain_kernels.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'k_anchor_targets|k_threshold_compact|k_segment_nms' -c 3 \
    -o $OUT/r2a_prof -f $KCMD > $OUT/r2a_ncu_full.log 2>&1
echo "ncufull_exit=$?"
tail -3 $OUT/r2a_ncu_full.log
