#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r2m_pytest.log 2>&1; echo "pytest_exit=$?"; grep -v "^frame" $OUT/r2m_pytest.log | tail -3
timeout 300 python profiles/nms_phases.py > $OUT/r2m_nms_phases.log 2>&1; head -13 $OUT/r2m_nms_phases.log
python __graft_entry__.py --smoke > $OUT/r2m_smoke.log 2>&1; echo "smoke_exit=$?"; tail -2 $OUT/r2m_smoke.log
timeout 900 python bench.py > $OUT/r2m_bench.json 2> $OUT/r2m_bench.err
echo "bench_exit=$?"; tail -2 $OUT/r2m_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2m_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, d["in_order"]["pages_per_s"])
for k in ("roofline", "roofline_k2", "roofline_k3"):
    print(k, d[k]["us_per_launch"], round(d[k]["frac"], 3))
print("nms", d["nms"]["us_per_launch"], d["nms"]["merge_us"], "inference", d["inference"]["reference_semantics"]["pages_per_s"])
for k in ("config3", "config4"):
    print(k, d[k]["pages_per_s"], d[k]["kernels_us"], {a: round(b, 3) for a, b in d[k]["hbm_fraction"].items() if a != "bytes"})
PY
