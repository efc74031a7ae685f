"""e2e step time of TargetLossStep.run_from_host for several chunk counts (one process, same box)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import retinanet_b200 as rn, synthetic
HW, B = (800, 1333), 16
anchors = rn.anchors_for_shape(HW + (3,))
N = anchors.shape[0]
images, anns = synthetic.training_batch(2, batch=B, anchors=np.asarray(anchors))
cls, reg = synthetic.training_predictions(2, B, N, classes=1)
cls_h, reg_h = torch.from_numpy(cls).pin_memory(), torch.from_numpy(reg).pin_memory()
step = rn.pipeline.TargetLossStep(HW + (3,), B, 22, 1)
import time
for gather in (False, True):
    for chunks in (1, 2, 4, 8):
        for _ in range(5):
            step.run_from_host(images, anns, cls_h, reg_h, chunks=chunks, gather_reg_from_host=gather)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            step.run_from_host(images, anns, cls_h, reg_h, chunks=chunks, gather_reg_from_host=gather)
        e1.record(); torch.cuda.synchronize()
        print("gather %d chunks %2d: %.3f ms/step" % (gather, chunks, e0.elapsed_time(e1) / 50))
t0 = time.perf_counter()
for _ in range(200):
    rn.anchors.pack_annotations(images, anns, 1)
print("pack_annotations: %.1f us" % ((time.perf_counter() - t0) / 200 * 1e6))
t0 = time.perf_counter()
for _ in range(200):
    step.load_annotations(images, anns)
torch.cuda.synchronize()
print("load_annotations: %.1f us" % ((time.perf_counter() - t0) / 200 * 1e6))
