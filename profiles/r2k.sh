#!/bin/bash
# round 2, capture k: select-based merge (C > 1), N4 conflict-free blur, K1 page-cost calibration of the new kernel
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/r2k_pytest.log 2>&1
echo "pytest_exit=$?" | tee -a $OUT/r2k_pytest.log
grep -v "^frame" $OUT/r2k_pytest.log | tail -12
timeout 600 python profiles/k1_page_cost.py > $OUT/r2k_k1_page_cost.log 2>&1
tail -4 $OUT/r2k_k1_page_cost.log
timeout 600 python - > $OUT/r2k_preprocess_time.log 2>&1 <<'PY'
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import retinanet_b200 as rn
from tests.test_oracle_preprocess import document_page, cv2_pipeline
B, H, W = 16, 2200, 1712
imgs = np.stack([document_page(100 + i, H, W) for i in range(4)] * 4)
dev = torch.from_numpy(imgs).cuda()
out = torch.empty_like(dev)
for _ in range(3):
    rn.preprocess.preprocess_pages(dev, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    rn.preprocess.preprocess_pages(dev, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
t0 = time.perf_counter(); cv2_pipeline(imgs[0]); t1 = time.perf_counter()
print("N4: %d pages of %dx%d in %.3f ms = %.0f pages/s; bytes 3 in + 3 out per pixel -> %.0f GB/s algorithmic; OpenCV one page on one core %.1f ms"
      % (B, H, W, ms, B / ms * 1e3, B * H * W * 6 / ms / 1e6, (t1 - t0) * 1e3))
PY
cat $OUT/r2k_preprocess_time.log | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/r2k_preprocess_launches.csv python - > /dev/null 2>&1 <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
import retinanet_b200 as rn
from tests.test_oracle_preprocess import document_page
imgs = np.stack([document_page(100 + i, 2200, 1712) for i in range(4)] * 4)
dev = torch.from_numpy(imgs).cuda()
for _ in range(2):
    rn.preprocess.preprocess_pages(dev)
torch.cuda.synchronize()
PY
grep -i "k_gray\|k_row\|k_distance" $OUT/r2k_preprocess_launches.csv | awk -F'","' '{print $1, $NF}' | tail -3
timeout 900 python bench.py --no-cpu > $OUT/r2k_bench.json 2> $OUT/r2k_bench.err
echo "bench_exit=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2k_bench.json"))
    print({k: d[k] for k in ("value", "ms_per_step")}, "K1", d["roofline"]["us_per_launch"], "K2", d["roofline_k2"]["us_per_launch"], "K3", d["roofline_k3"]["us_per_launch"], d["roofline_k3"]["frac"], "NMS", d["nms"]["us_per_launch"])
    print("inference", d["inference"]["reference_semantics"]["pages_per_s"], d["inference"]["reference_semantics"]["ms_per_batch"])
    for k in ("config3", "config4"):
        print(k, d[k]["pages_per_s"], d[k]["kernels_us"])
except Exception as e:
    print("bench parse failed", repr(e))
PY
