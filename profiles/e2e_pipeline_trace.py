"""Device-side timeline of the pipelined host-input step (HostStepPipeline): when do K1 / K2 of each step start and
end relative to the copies?  Events are recorded around the K1 graph replay and the K2 launch (after its wait for the
copy), on the pipeline's compute stream.

    PROBE_DELAY_US=0|220 python profiles/e2e_pipeline_trace.py
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
DELAY = float(os.environ.get("PROBE_DELAY_US", "0")) * 1e-6
GATHER = os.environ.get("PROBE_GATHER", "1") == "1"
depth = int(os.environ.get("PROBE_DEPTH", "2"))
pipe = rn.pipeline.HostStepPipeline(HW + (3,), B, 22, 1, depth=depth)
trace = []


class Replay(object):
    def __init__(self, g):
        self.g = g

    def replay(self):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        self.g.replay()
        b.record()
        trace.append(["k1", a, b])


for s in pipe.slots:
    s._graphs = (Replay(s._graphs[0]), s._graphs[1])
orig = rn.pipeline._losses.detection_losses


def traced(*a, **k):
    x, y = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    x.record()
    r = orig(*a, **k)
    y.record()
    trace.append(["k2", x, y])
    return r


rn.pipeline._losses.detection_losses = traced
t_origin = torch.cuda.Event(enable_timing=True)
pend = []
for i in range(40):
    if i == 20:
        del trace[:]
        t_origin.record()
    a = time.perf_counter()
    pend.append(pipe.submit(images, anns, cls_h, reg_h, chunks=1, gather_reg_from_host=GATHER))
    while DELAY and time.perf_counter() - a < DELAY:
        pass
    if len(pend) == depth:
        pipe.result(pend.pop(0))
pipe.drain()
torch.cuda.synchronize()
print("delay %.0f us, depth %d, gather %d" % (DELAY * 1e6, depth, GATHER))
for name, x, y in trace[:24]:
    print("%s  start %8.1f us  end %8.1f us  (%.1f us)" % (name, t_origin.elapsed_time(x) * 1e3, t_origin.elapsed_time(y) * 1e3, x.elapsed_time(y) * 1e3))
