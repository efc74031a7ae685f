#!/bin/bash
# round 2, capture l (8 GPUs): multi-GPU checks on 8 ranks, bench at N = 8 and N = 4
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 profiles/multigpu_check.py > $OUT/r2l_multigpu.log 2>&1
echo "multigpu_exit=$?"; grep -v "^frame\|^W1\|^\*\*\*\|^\[W" $OUT/r2l_multigpu.log | tail -6
for N in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 30 --no-cpu > $OUT/r2l_bench_n$N.json 2> $OUT/r2l_bench_n$N.err
echo "bench_n${N}_exit=$?"; tail -2 $OUT/r2l_bench_n$N.err
python - $N <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/r2l_bench_n%s.json" % sys.argv[1]))
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, "K1 by rank", [round(x, 1) for x in d["roofline"]["us_per_launch_by_rank"]], "K2", d["roofline_k2"]["us_per_launch"])
    print("in_order", d["in_order"]["pages_per_s"], "e2e", d["e2e"]["value"], d["e2e"]["h2d_ceiling_GBps"], d["e2e"]["fraction_of_h2d_ceiling"], d["losses"])
    print("inference", d["inference"]["reference_semantics"]["pages_per_s"], d["inference"]["reference_semantics"]["e2e_pages_per_s"], "c3", d["config3"]["pages_per_s"], "c4", d["config4"]["pages_per_s"])
except Exception as e:
    print("bench parse failed", repr(e))
PY
done
