"""K1's time as a function of the page: 16 copies of one page per launch, for 32 pages of BASELINE config 2, next to
the overlap estimate distributed.page_cost() uses for load-aware sharding (fit: time ~ a + b * pairs).

    python profiles/k1_page_cost.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import retinanet_b200 as rn  # noqa: E402
import synthetic  # noqa: E402

HW, B = (800, 1333), 16
anchors = rn.anchors_for_shape(HW + (3,))
images, anns = synthetic.training_batch(2, batch=32, anchors=np.asarray(anchors))
step = rn.pipeline.TargetLossStep(HW + (3,), B, 22, 1)
rows = []
for i, a in enumerate(anns):
    step.load_annotations(images[:B], [a] * B)
    step._graphs or step._build_graphs()
    for _ in range(3):
        step._graphs[0].replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        step._graphs[0].replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    bb = np.asarray(a['bboxes'])
    area = float(((bb[:, 2] - bb[:, 0]) * (bb[:, 3] - bb[:, 1])).sum()) / (HW[0] * HW[1])
    pairs = rn.distributed.page_cost([a], HW, fixed=0.0)[0]
    rows.append((len(bb), area, pairs, us))
    print("page %2d  G %2d  area %.3f  pairs %.0f  K1 %.2f us" % (i, len(bb), area, pairs, us))
r = np.array(rows)
for name, col in (("G", 0), ("area", 1), ("pairs", 2)):
    A = np.stack([np.ones(len(r)), r[:, col]], 1)
    coef, res, _, _ = np.linalg.lstsq(A, r[:, 3], rcond=None)
    pred = A @ coef
    print("fit on %-5s: us = %.2f + %.4g * x   rms error %.2f us (of mean %.1f)" % (name, coef[0], coef[1], np.sqrt(np.mean((pred - r[:, 3]) ** 2)), r[:, 3].mean()))
A = np.stack([np.ones(len(r)), r[:, 0], r[:, 2]], 1)
coef, res, _, _ = np.linalg.lstsq(A, r[:, 3], rcond=None)
pred = A @ coef
print("fit on G + pairs: us = %.2f + %.4g * G + %.4g * pairs   rms error %.2f us" % (coef[0], coef[1], coef[2], np.sqrt(np.mean((pred - r[:, 3]) ** 2))))
