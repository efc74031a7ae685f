"""The detection tail at BASELINE configs[4] (16 pages of 800x1333, 80 classes) launched eagerly a few times, for ncu captures
of k_threshold_keys_stream / k_segment_nms / k_merge_topk_select at C = 80.

    python profiles/run_kernels_c80.py [reps]
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
hw, C, B = (800, 1333), 80, 16
anchors = np.asarray(rn.anchors_for_shape(hw + (3,)))
images, anns = synthetic.training_batch(5, batch=B, anchors=anchors)
cls_d, reg_d = synthetic.inference_predictions_torch(5, B, anchors, anns, classes=C, device=torch.device("cuda"))
det = rn.pipeline.DetectionStep(hw, B, C, use_graph=False)
det.cls_pred, det.reg_pred = cls_d, reg_d
for _ in range(reps):
    det.run()
torch.cuda.synchronize()
print("detections", int((det.scores >= 0).sum()))
