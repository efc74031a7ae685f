#!/bin/bash
# round 2, capture g (2 GPUs): multi-GPU checks (mailbox loss exchange, no-mailbox pipelined schedule, timeout, two devices in
# one process) and the bench line at N = 2
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python profiles/multigpu_check.py --two-devices > $OUT/r2g_two_devices.log 2>&1
echo "two_devices_exit=$?"; tail -3 $OUT/r2g_two_devices.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 profiles/multigpu_check.py > $OUT/r2g_multigpu.log 2>&1
echo "multigpu_exit=$?"; grep -v "^frame\|^W1\|^\*\*\*" $OUT/r2g_multigpu.log | tail -25
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 > $OUT/r2g_bench_n2.json 2> $OUT/r2g_bench_n2.err
echo "bench_n2_exit=$?"; tail -3 $OUT/r2g_bench_n2.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2g_bench_n2.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, "K1", d["roofline"]["us_per_launch_by_rank"], "K2", d["roofline_k2"]["us_per_launch"])
    print("in_order", d["in_order"], "e2e", d["e2e"]["value"], d["e2e"]["h2d_ceiling_GBps"], d["e2e"]["fraction_of_h2d_ceiling"], d["losses"])
    print("inference", d["inference"]["reference_semantics"]["pages_per_s"], "c3", d["config3"]["pages_per_s"], "c4", d["config4"]["pages_per_s"])
except Exception as e:
    print("bench parse failed", repr(e))
PY
