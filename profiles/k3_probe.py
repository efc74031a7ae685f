"""K3 at the benchmark's inference sizes with (a) the real threshold, (b) a threshold nothing passes: the
difference is what the candidate phase (gather + decode + slab stores) costs on top of streaming the scores."""
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
for thr in (0.05, 2.0):
    head = rn.DetectionHead(score_threshold=thr)
    for _ in range(3):
        head([shape, reg_d, cls_d])
torch.cuda.synchronize()
