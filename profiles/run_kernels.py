"""Launch each hot-path kernel a few times at the benchmark's sizes (for ncu captures).

    python profiles/run_kernels.py [reps]

K1 + K2 at BASELINE configs[1] (16 pages, 800x1333, C=1), K3 + K4/K5 + merge at configs[2] (64 pages).
No CUDA graphs here so every launch is a plain kernel node for the profiler.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import retinanet_b200 as rn  # noqa: E402
import synthetic  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
HW = (800, 1333)
anchors = rn.anchors_for_shape(HW + (3,))
N = anchors.shape[0]
B = 16
images, anns = synthetic.training_batch(2, batch=B, anchors=np.asarray(anchors))
cls, reg = synthetic.training_predictions(2, B, N, classes=1)
step = rn.pipeline.TargetLossStep(HW + (3,), B, 22, 1, use_graph=False)
step.load_annotations(images, anns)
step.load_predictions(torch.from_numpy(cls), torch.from_numpy(reg))
for _ in range(reps):
    step.run()
torch.cuda.synchronize()
print("losses", step.losses.cpu().numpy())

Bi = 64
_, anns_i = synthetic.training_batch(3, batch=Bi)
icls, ireg = synthetic.inference_predictions(3, Bi, np.asarray(anchors), anns_i, classes=1)
icls_d, ireg_d = torch.from_numpy(icls).cuda(), torch.from_numpy(ireg).cuda()
head = rn.DetectionHead()
for _ in range(reps):
    out = head([(Bi,) + HW + (3,), ireg_d, icls_d])
torch.cuda.synchronize()
print("detections", int((out[1] >= 0).sum()))
