"""Per-kernel device times of the inference path at the benchmark's sizes (64 pages, 800x1333, C=1):
K3 alone (score threshold so high that NMS has nothing to do is NOT used -- instead the three kernels are timed
through CUDA events around whole calls with nms on/off and max_detections small).  Prints mean us per batch."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import retinanet_b200 as rn  # noqa: E402
import synthetic  # noqa: E402

HW, B = (800, 1333), 64
anchors = np.asarray(rn.anchors_for_shape(HW + (3,)))
_, anns = synthetic.training_batch(3, batch=B)
cls, reg = synthetic.inference_predictions(3, B, anchors, anns, classes=1)
cls_d, reg_d = torch.from_numpy(cls).cuda(), torch.from_numpy(reg).cuda()
shape = (B,) + HW + (3,)


def timed(head, n=30):
    for _ in range(3):
        head([shape, reg_d, cls_d])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        head([shape, reg_d, cls_d])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


full = timed(rn.DetectionHead())
top1k = timed(rn.DetectionHead(pre_nms_top_k=1000))
# max_detections=1 ends k_segment_nms after the first selected box: what is left is K3 + one radix select/sort round
k3ish = timed(rn.DetectionHead(max_detections=1, pre_nms_top_k=1))
print("full %.1f us   pre_nms_top_k=1000 %.1f us   K3 + minimal back end %.1f us" % (full, top1k, k3ish))
