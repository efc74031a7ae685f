"""Per-CTA trace of K1 (k_anchor_targets_tiles32) at the benchmark's sizes: start / end (%globaltimer), SM, level, page of
every CTA of ONE launch, from the A/B build with -DRN_K1_TRACE (python retinanet-for-table-detection_b200/build.py
--variant trace RN_K1_TRACE=1).  Prints the mean CTA duration per level, the launch's span, the busy share of the CTA
slots, what the last 10 us are made of, and writes the raw table to gpurun_out/k1_cta_trace.npy for offline what-if
scheduling (profiles/k1_order_sim.py).

    RN_B200_LIB=retinanet-for-table-detection_b200/librn_b200.trace.so python profiles/k1_cta_trace.py
"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import retinanet_b200 as rn
    import synthetic
    lib = rn._lib.load()
    fn = lib.rn_debug_k1_trace
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
    HW, B, gmax = (800, 1333), 16, 22
    anchors = rn.anchors_for_shape(HW + (3,))
    images, anns = synthetic.training_batch(2, batch=B, anchors=np.asarray(anchors))
    step = rn.pipeline.TargetLossStep(HW + (3,), B, gmax, 1)
    step.load_annotations(images, anns)
    step._build_graphs()
    for _ in range(20):
        step._graphs[0].replay()
    torch.cuda.synchronize()
    out = np.zeros((16384, 4), dtype=np.uint64)
    assert fn(out.ctypes.data, 16384) == 0
    out = out[out[:, 0] > 0]                               # the CTAs of the launch (the trace buffer starts out zeroed)
    n = len(out)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.save(os.path.join(ROOT, "gpurun_out", "k1_cta_trace.npy"), out)
    t0, t1 = out[:, 0].astype(np.int64), out[:, 1].astype(np.int64)
    sm = out[:, 2].astype(np.int64)
    level = (out[:, 3] & np.uint64(0xff)).astype(np.int64)
    page = ((out[:, 3] >> np.uint64(8)) & np.uint64(0xff)).astype(np.int64)
    base = t0.min()
    t0, t1 = (t0 - base) / 1e3, (t1 - base) / 1e3
    dur = t1 - t0
    print("CTAs %d  span %.2f us  sum of CTA durations %.1f us = %.3f of 444 slots x span" % (n, t1.max(), dur.sum(), dur.sum() / (444 * t1.max())))
    for l in range(5):
        m = level == l
        print("level %d: %4d CTAs  mean %.2f us  min %.2f  max %.2f  share of CTA time %.3f" % (l, m.sum(), dur[m].mean(), dur[m].min(), dur[m].max(), dur[m].sum() / dur.sum()))
    for p in np.argsort([-dur[page == q].sum() for q in range(B)]):
        m = page == p
        print("page %2d: CTA time %.1f us  first start %.2f  last end %.2f" % (p, dur[m].sum(), t0[m].min(), t1[m].max()))
    end = t1.max()
    for w in (15, 10, 5, 2):
        m = t1 > end - w
        busy = (np.minimum(t1[m], end) - np.maximum(t0[m], end - w)).sum() / (444 * w)
        print("last %2d us: %3d CTAs alive, slots busy %.2f, levels %s" % (w, m.sum(), busy, np.bincount(level[m], minlength=5).tolist()))
    per_sm = np.array([t1[sm == s].max() if (sm == s).any() else 0 for s in range(148)])
    print("per-SM last end: min %.2f mean %.2f max %.2f" % (per_sm.min(), per_sm.mean(), per_sm.max()))
    starts = np.sort(t0)
    print("start times: 444th CTA %.2f us, last CTA %.2f us" % (starts[min(443, n - 1)], starts[-1]))


if __name__ == "__main__":
    main()
