"""Where does the pipelined host-input step spend its time?  CPU seconds inside submit() (GT packing, enqueue) against
the wait inside result(); depth 2 / 3; gather on / off.

    python profiles/e2e_pipeline_probe.py
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import retinanet_b200 as rn  # noqa: E402
import synthetic  # noqa: E402

HW, B = (800, 1333), 16
anchors = rn.anchors_for_shape(HW + (3,))
N = anchors.shape[0]
images, anns = synthetic.training_batch(2, batch=B, anchors=np.asarray(anchors))
cls, reg = synthetic.training_predictions(2, B, N, classes=1)
cls_h, reg_h = torch.from_numpy(cls).pin_memory(), torch.from_numpy(reg).pin_memory()
# the box's H2D ceiling for this tensor and the synchronous step, for reference
dev_buf = torch.empty_like(cls_h, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    dev_buf.copy_(cls_h, non_blocking=True)
torch.cuda.synchronize()
e0.record()
for _ in range(50):
    dev_buf.copy_(cls_h, non_blocking=True)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 50 * 1e3
print("H2D of the classification tensor alone: %.1f us (%.1f GB/s) -> ceiling %.0f pages/s" % (us, cls_h.numel() * 4 / us / 1e3, B / us * 1e6))
sync = rn.pipeline.TargetLossStep(HW + (3,), B, 22, 1)
for timed in (False, True):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(100 if timed else 10):
        sync.run_from_host(images, anns, cls_h, reg_h, chunks=1, gather_reg_from_host=True)
    t1 = time.perf_counter()
print("synchronous run_from_host: %.1f us/step" % (1e6 * (t1 - t0) / 100))
del sync
DELAY = float(os.environ.get("PROBE_DELAY_US", "0")) * 1e-6
for depth, gather in ((2, True), (3, True), (4, True), (2, False)):
    pipe = rn.pipeline.HostStepPipeline(HW + (3,), B, 22, 1, depth=depth)
    for timed in (False, True):
        steps = 100 if timed else 10
        t_sub = t_res = 0.0
        pend = []
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            a = time.perf_counter()
            pend.append(pipe.submit(images, anns, cls_h, reg_h, chunks=1, gather_reg_from_host=gather))
            while DELAY and time.perf_counter() - a < DELAY:
                pass
            b = time.perf_counter()
            if len(pend) == depth:
                pipe.result(pend.pop(0))
            c = time.perf_counter()
            t_sub += b - a
            t_res += c - b
        pipe.drain()
        t1 = time.perf_counter()
    print("depth %d gather %d: %.1f us/step (%.0f pages/s)  submit %.1f us  result-wait %.1f us"
          % (depth, gather, 1e6 * (t1 - t0) / steps, B * steps / (t1 - t0), 1e6 * t_sub / steps, 1e6 * t_res / steps))
    del pipe
