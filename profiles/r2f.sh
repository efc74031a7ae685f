#!/bin/bash
# round 2, capture f: new bench.py (all four configurations), tests incl. full-batch configs, NMS window-by-cost, K3 sparse loop
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/r2f_pytest.log 2>&1
echo "pytest_exit=$?" | tee -a $OUT/r2f_pytest.log
grep -v "^frame" $OUT/r2f_pytest.log | tail -30
timeout 300 python profiles/nms_phases.py > $OUT/r2f_nms_phases.log 2>&1
cat $OUT/r2f_nms_phases.log
timeout 300 python profiles/time_inference.py > $OUT/r2f_time_inference.log 2>&1
cat $OUT/r2f_time_inference.log
timeout 900 python bench.py > $OUT/r2f_bench.json 2> $OUT/r2f_bench.err
echo "bench_exit=$?"
tail -5 $OUT/r2f_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2f_bench.json"))
    print({k: d[k] for k in ("value", "ms_per_step")}, "K1", d["roofline"]["us_per_launch"], d["roofline"]["frac"], "K2", d["roofline_k2"]["us_per_launch"], d["roofline_k2"]["frac"])
    print("K3", d["roofline_k3"]["us_per_launch"], d["roofline_k3"]["frac"], "NMS", d["nms"]["us_per_launch"])
    print("e2e", d["e2e"]["value"], d["e2e"]["fraction_of_h2d_ceiling"])
    print("inference", {k: v for k, v in d["inference"]["reference_semantics"].items() if not isinstance(v, (dict, str))})
    for k in ("config3", "config4"):
        print(k, d[k]["pages_per_s"], d[k]["kernels_us"], d[k]["hbm_fraction"])
    print("cpu", d["cpu_baseline"]["value"], d["inference"]["cpu_baseline"]["value"], d["clocks"])
except Exception as e:
    print("bench parse failed", repr(e))
PY
KCMD="python profiles/run_kernels.py 1"
timeout 300 $KCMD > $OUT/r2f_plain_kernels.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'k_threshold_keys|k_segment_nms' -c 2 \
    -o $OUT/r2f_prof -f $KCMD > $OUT/r2f_ncu_full.log 2>&1
echo "ncufull_exit=$?"
