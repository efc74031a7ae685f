"""Generate the committed golden fixtures in this directory.

Run from the repo root, in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

* ``numpy_half.npz``  -- produced by executing the REFERENCE'S OWN CODE (``model/anchors.py``,
  ``model/utils.py:compute_overlap``) imported unmodified through ``oracle/ref_loader.py``.
  These pin the oracle (and the CUDA path) to the reference bit-for-bit.
* ``tf_half.npz``     -- produced by the oracle's restatement of the TensorFlow half
  (``oracle/losses_np.py``, ``oracle/layers_np.py``).  The reference cannot be executed for
  this half (no TensorFlow), so these are regression fixtures for the restatement
  ("parity unpinned", see ``oracle/__init__.py``), not reference outputs.

Large arrays are stored as sha256 digests plus the sparse parts (positive / ignore indices,
rows at positives); small cases are stored in full.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import synthetic  # noqa: E402
from oracle import layers_np, losses_np  # noqa: E402
from oracle.ref_loader import load_reference  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def tie_cases():
    """Explicit-anchor cases probing the threshold / tie semantics."""
    cases = {}
    # IoU exactly 0.5 -> positive; exactly 0.4 -> background (strict > in fp32); 0.25
    cases['exact'] = (np.array([[0, 0, 32, 32], [0, 0, 32, 32], [100, 100, 132, 132], [0, 0, 10, 10]], dtype=np.float64),
                      np.array([[0, 0, 32, 16], [200, 200, 232, 212.8], [100, 100, 132, 112.8]], dtype=np.float64),
                      np.array([0., 1., 2.]))
    # duplicate GT boxes, equal IoU for several GT -> lowest GT index
    cases['dup'] = (np.array([[10, 10, 50, 50], [12, 8, 48, 52], [300, 300, 340, 340]], dtype=np.float64),
                    np.array([[10, 10, 50, 50], [10, 10, 50, 50], [11, 9, 49, 51], [11, 9, 49, 51]], dtype=np.float64),
                    np.array([2., 1., 0., 1.]))
    # ties that only appear after rounding the fp64 IoU to fp32
    a = np.array([[0, 0, 100, 100]], dtype=np.float64)
    g = np.array([[0, 0, 100, 70.0000000001], [0, 0, 100, 70.0]], dtype=np.float64)
    cases['f32tie'] = (a, g, np.array([1., 0.]))
    # degenerate GT (zero area) and a GT far outside
    cases['degenerate'] = (np.array([[0, 0, 16, 16], [8, 8, 24, 24]], dtype=np.float64),
                           np.array([[4, 4, 4, 12], [1000, 1000, 1100, 1100]], dtype=np.float64),
                           np.array([0., 0.]))
    return cases


def main():
    ref_anchors, ref_utils = load_reference()
    out = {}

    # ---- A1/A2: anchor generation -------------------------------------------------------
    for size in (32, 64, 128, 256, 512):
        out['base_%d' % size] = ref_anchors.generate_anchors(size)
    a800 = ref_anchors.anchors_for_shape((800, 1333, 3))
    rs = np.random.RandomState(7)
    rows = np.sort(np.concatenate([np.arange(20), a800.shape[0] - 1 - np.arange(20),
                                   rs.randint(0, a800.shape[0], 200)]))
    out['a800_shape'] = np.array(a800.shape)
    out['a800_sha'] = sha(a800)
    out['a800_rows_idx'] = rows
    out['a800_rows'] = a800[rows]
    a1600 = ref_anchors.anchors_for_shape((1600, 2400, 3))
    out['a1600_shape'] = np.array(a1600.shape)
    out['a1600_sha'] = sha(a1600)
    custom = ref_anchors.AnchorParameters([24, 48, 100, 200, 400], [8, 16, 32, 64, 128],
                                          np.array([0.3, 1, 2.5], np.float32), np.array([1, 1.3], np.float32))
    out['custom_anchors'] = ref_anchors.anchors_for_shape((70, 90, 3), anchor_params=custom)
    out['small_anchors'] = ref_anchors.anchors_for_shape((67, 93, 3))

    # ---- A3: compute_overlap on a small random case (full matrix) --------------------------
    rs = np.random.RandomState(11)
    sm = out['small_anchors']
    gsm = np.stack([rs.uniform(0, 40, 6), rs.uniform(0, 30, 6), rs.uniform(45, 93, 6), rs.uniform(32, 67, 6)], axis=1)
    out['overlap_gt'] = gsm
    out['overlap_small'] = ref_utils.compute_overlap(sm, gsm)

    # ---- A5: small batch, stored in full (C=3, mixed page widths, one empty page) ---------
    imgs = [synthetic.PageShape((67, 93, 3)), synthetic.PageShape((60, 80, 3)), synthetic.PageShape((67, 93, 3))]
    anns = [{'bboxes': gsm[:4], 'labels': np.array([0., 2., 1., 1.])},
            {'bboxes': gsm[3:], 'labels': np.array([1., 0., 2.])},
            {'bboxes': np.zeros((0, 4)), 'labels': np.zeros((0,))}]
    reg, lab = ref_anchors.anchor_targets_bbox(sm, imgs, anns, 3)
    out['small_reg'] = reg
    out['small_lab'] = lab
    for i, a in enumerate(anns):
        out['small_gt_%d' % i] = a['bboxes']
        out['small_gl_%d' % i] = a['labels']
    out['small_shapes'] = np.array([im.shape for im in imgs])

    # ---- A5: the survey's 3-GT page and the empty page at 800x1333 -----------------------------
    page3 = {'bboxes': np.array([[100, 100, 600, 400], [50, 500, 700, 760], [800, 100, 1300, 300]], dtype=np.float64),
             'labels': np.zeros(3)}
    reg, lab = ref_anchors.anchor_targets_bbox(a800, [synthetic.PageShape((800, 1333, 3))], [page3], 1)
    out['page3_pos'] = np.nonzero(lab[0, :, -1] == 1)[0]
    out['page3_ign'] = np.nonzero(lab[0, :, -1] == -1)[0]
    out['page3_reg_sha'] = sha(reg)
    out['page3_lab_sha'] = sha(lab)
    out['page3_reg_pos_rows'] = reg[0, out['page3_pos']]
    _, _, amax = ref_anchors.compute_gt_annotations(a800, page3['bboxes'])
    out['page3_argmax_sha'] = sha(amax.astype(np.int32))
    empty = {'bboxes': np.zeros((0, 4)), 'labels': np.zeros((0,))}
    reg, lab = ref_anchors.anchor_targets_bbox(a800, [synthetic.PageShape((800, 1333, 3))], [empty], 1)
    out['empty_ign'] = np.nonzero(lab[0, :, -1] == -1)[0]
    out['empty_reg_sha'] = sha(reg)
    out['empty_lab_sha'] = sha(lab)

    # ---- tie / threshold cases with explicit anchors (stored in full) ----------------------------
    for name, (anc, gt, gl) in tie_cases().items():
        reg, lab = ref_anchors.anchor_targets_bbox(anc, [synthetic.PageShape((2000, 2000, 3))],
                                                   [{'bboxes': gt, 'labels': gl}], 3)
        pos, ign, amax = ref_anchors.compute_gt_annotations(anc, gt)
        out['tie_%s_anchors' % name] = anc
        out['tie_%s_gt' % name] = gt
        out['tie_%s_gl' % name] = gl
        out['tie_%s_reg' % name] = reg
        out['tie_%s_lab' % name] = lab
        out['tie_%s_argmax' % name] = amax

    # ---- seeded synthetic pages of the BASELINE configs (digests + sparse parts) ---------------------
    for cfg, batch, mixed in ((1, 1, False), (2, 3, True), (5, 1, False), (4, 1, False)):
        c = synthetic.CONFIGS[cfg]
        anc = a800 if c['hw'] == (800, 1333) else a1600
        imgs, anns = synthetic.training_batch(cfg, batch=batch, mixed_widths=mixed, anchors=anc)
        reg, lab = ref_anchors.anchor_targets_bbox(anc, imgs, anns, c['classes'])
        st = lab[:, :, -1]
        out['cfg%d_reg_sha' % cfg] = sha(reg)
        out['cfg%d_lab_sha' % cfg] = sha(lab)
        out['cfg%d_state_sha' % cfg] = sha(st)
        out['cfg%d_npos' % cfg] = np.array([(st[b] == 1).sum() for b in range(batch)])
        out['cfg%d_nign' % cfg] = np.array([(st[b] == -1).sum() for b in range(batch)])
        out['cfg%d_pos0' % cfg] = np.nonzero(st[0] == 1)[0]
        out['cfg%d_reg_pos0' % cfg] = reg[0, out['cfg%d_pos0' % cfg]]

    np.savez_compressed(os.path.join(HERE, 'numpy_half.npz'), **out)
    print('numpy_half.npz: %d entries' % len(out))

    # ================= TF half: restatement fixtures (unpinned) ==================================
    tf = {}
    rs = np.random.RandomState(21)
    B, N, C = 2, 400, 3
    y_cls = np.zeros((B, N, C + 1), np.float32)
    state = rs.choice([-1, 0, 1], (B, N), p=[0.1, 0.75, 0.15]).astype(np.float32)
    y_cls[:, :, -1] = state
    hot = rs.randint(0, C, (B, N))
    for b in range(B):
        for n in np.nonzero(state[b] == 1)[0]:
            y_cls[b, n, hot[b, n]] = 1
    p = (1 / (1 + np.exp(-rs.normal(-2, 2, (B, N, C))))).astype(np.float32)
    p[0, 0, 0], p[0, 1, 0], p[0, 2, 1], p[0, 3, 2] = 0.0, 1.0, 1e-8, 1 - 1e-8
    y_reg = np.concatenate([rs.normal(0, 1, (B, N, 4)).astype(np.float32), state[:, :, None]], axis=2)
    r = (y_reg[:, :, :4] + rs.normal(0, 0.2, (B, N, 4)).astype(np.float32)).astype(np.float32)
    r[0, 5] = y_reg[0, 5, :4]                 # zero diff -> sign(0) = 0
    tf['loss_y_cls'], tf['loss_p'], tf['loss_y_reg'], tf['loss_r'] = y_cls, p, y_reg, r
    for mode in ('tf2', 'logits'):
        l, g = losses_np.focal(bce=mode)(y_cls, p, return_grad=True)
        tf['focal_%s_loss' % mode], tf['focal_%s_grad' % mode] = np.float32(l), g
    l, g = losses_np.smooth_l1()(y_reg, r, return_grad=True)
    tf['sl1_loss'], tf['sl1_grad'] = np.float32(l), g

    hw = (67, 93)
    anc32 = layers_np.all_anchors_f32(hw, batch=1)
    tf['det_anchors_f32'] = anc32[0]
    n = anc32.shape[1]
    rs = np.random.RandomState(22)
    reg = rs.normal(0, 0.5, (2, n, 4)).astype(np.float32)
    cls = (1 / (1 + np.exp(-rs.normal(-2.5, 1.5, (2, n, 3))))).astype(np.float32)
    cls[0, 10:40, 1] = cls[0, 10, 1]            # equal scores -> index order
    tf['det_reg'], tf['det_cls'] = reg, cls
    boxes = layers_np.bbox_transform_inv(np.tile(anc32, (2, 1, 1)), reg)
    tf['det_boxes'] = boxes
    tf['det_clipped'] = layers_np.clip_boxes(hw, boxes)
    for tag, kw in (('default', {}), ('agnostic', dict(class_specific_filter=False)),
                    ('nonms', dict(nms=False)), ('small', dict(max_detections=20, nms_threshold=0.3, score_threshold=0.2))):
        res = layers_np.detect(hw, reg, cls, **kw)
        tf['det_%s_boxes' % tag], tf['det_%s_scores' % tag], tf['det_%s_labels' % tag], tf['det_%s_idx' % tag] = res
    np.savez_compressed(os.path.join(HERE, 'tf_half.npz'), **tf)
    print('tf_half.npz: %d entries' % len(tf))


if __name__ == '__main__':
    main()
