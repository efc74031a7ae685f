"""Per-phase clock64 breakdown of k_segment_nms (rn_debug_nms_timing) at the benchmark's inference sizes."""
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
names = [  # phases end at block barriers: thread 0's clock is the block's
         "bisection select", "gather", "sort (warp bitonic + merge levels)", "group start (normalise, first window)", "pairs (conflict matrix of the batch)", "resolve (warp 0)", "sweep between rounds", "rank the alive (per batch/window)", "apply (new selections vs open windows + next window vs all)", "-", "chunk: read + decode all rows", "group: closing barrier"]
rn._lib.load().rn_debug_nms_timing(1)
for topk in (0, 1000):
    head = rn.DetectionHead(pre_nms_top_k=topk)
    for _ in range(3):
        head([(B,) + HW + (3,), reg_d, cls_d])
    torch.cuda.synchronize()
    ws = [v for k, v in rn._lib._scratch.items() if k[0] == "filter"][0]
    full = ws[:2048].view(torch.int64).cpu().numpy()
    raw = ws[:256].view(torch.int64).cpu().numpy().astype(np.float64)
    slowest, t = raw[7], np.concatenate([raw[:7], raw[8:10], raw[17:20]])
    print("pre_nms_top_k=%d: ticks per segment (thread 0): mean %.0f, slowest CTA %.0f" % (topk, t.sum() / B, slowest))
    for n, v in zip(names, t):
        print("   %-30s %8.0f  %5.1f%%" % (n, v / B, 100 * v / t.sum()))
    print("   per page: %.1f batches, %.1f windows (%.0f candidates opened), %.0f candidates sorted in %.2f rounds, %.0f selected, %.0f above the threshold"
          % tuple(raw[k] / B for k in (10, 11, 16, 12, 14, 13, 15)))
    if topk == 0:
        print("   first 56 pages: ticks, above threshold, sorted, opened, rounds, batches, windows")
        rec = full[32:].reshape(56, 4)
        for i in np.argsort(-rec[:, 0]):
            r = rec[i]
            print("   page %2d %7d  %5d %5d %5d  %d  %3d %3d" % (i, r[0], r[1] & 0xffffffff, r[1] >> 32, r[2] & 0xffffffff, r[2] >> 32,
                                                                 r[3] & 0xffffffff, r[3] >> 32))
