"""Generate tests/golden/generator_half.npz from the reference's OWN ``csv_generator.Generator`` methods
(imported unmodified under stub keras / tensorflow modules, oracle/ref_loader.py).  Run in the build container,
where /root/reference exists:

    python tests/golden/make_golden_generator.py

Stored: inputs + outputs of ``filter_annotations`` (boxes on / across every validity edge), ``compute_inputs``
(ragged page sizes) and ``compute_targets`` (mixed page sizes -> anchors of the batch-max shape; each page's own
shape drives the border-ignore rule).
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402


class Page(object):
    def __init__(self, shape):
        self.shape = tuple(shape)


def cases():
    out = {}
    # ---- filter_annotations: every edge of the six rules, per image shape (H, W) ----
    shapes = np.array([[100, 200, 3], [64, 48, 3], [37, 91, 1]])
    boxes, labels = [], []
    for h, w, _ in shapes:
        b = np.array([
            [0, 0, 10, 10], [5, 5, 5, 10], [5, 5, 10, 5], [6, 5, 5, 10], [-1e-9, 0, 4, 4], [0, -0.5, 4, 4],
            [0, 0, w, h], [0, 0, w + 1e-9, h], [0, 0, w, h + 0.25], [w - 1, h - 1, w, h], [1.5, 2.5, 3.5, 4.5],
            [0, 0, 0, 0], [w, h, w, h], [3, 3, w + 5, h + 5]], np.float64)
        boxes.append(b)
        labels.append(np.arange(len(b), dtype=np.float64))
    out['flt_shapes'] = shapes
    out['flt_boxes'] = np.stack(boxes)
    out['flt_labels'] = np.stack(labels)
    # ---- compute_inputs ----
    rs = np.random.RandomState(5)
    out['inp_shapes'] = np.array([[5, 7, 3], [8, 4, 3], [6, 6, 3]])
    for i, s in enumerate(out['inp_shapes']):
        out['inp_img%d' % i] = rs.uniform(-1, 1, tuple(s)).astype(np.float32)
    # ---- compute_targets: mixed page sizes ----
    out['tgt_shapes'] = np.array([[128, 160, 3], [120, 200, 3], [96, 96, 3]])
    gts = [np.array([[10, 12, 90, 70], [60, 40, 150, 110.5]]), np.array([[5, 5, 190, 60], [100, 70, 180, 118], [20, 80, 70, 115]]),
           np.zeros((0, 4))]
    for i, g in enumerate(gts):
        out['tgt_boxes%d' % i] = g.astype(np.float64)
        out['tgt_labels%d' % i] = (np.arange(len(g)) % 2).astype(np.float64)
    return out


def main():
    gen_mod = ref_loader.load_reference_generator()
    G = gen_mod.Generator
    g = G.__new__(G)
    data = cases()
    # filter_annotations
    pages = [Page(s) for s in data['flt_shapes']]
    anns = [{'bboxes': data['flt_boxes'][i].copy(), 'labels': data['flt_labels'][i].copy()} for i in range(len(pages))]
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        _, kept = g.filter_annotations(pages, anns, list(range(len(pages))))
    data['flt_warnings'] = np.array(len(caught))
    for i, a in enumerate(kept):
        data['flt_out_boxes%d' % i] = a['bboxes']
        data['flt_out_labels%d' % i] = a['labels']
    # compute_inputs
    imgs = [data['inp_img%d' % i] for i in range(len(data['inp_shapes']))]
    g.batch_size = len(imgs)
    data['inp_out'] = g.compute_inputs(imgs)
    # compute_targets (2 classes; default anchor parameters, guess_shapes, overlaps from model/Parameters.py)
    g.load_parm_from_config = False
    g.compute_shapes = gen_mod.guess_shapes
    g.compute_anchor_targets = gen_mod.anchor_targets_bbox
    g.num_classes = lambda: 2
    pages = [Page(s) for s in data['tgt_shapes']]
    anns = [{'bboxes': data['tgt_boxes%d' % i], 'labels': data['tgt_labels%d' % i]} for i in range(len(pages))]
    reg, lab = g.compute_targets(pages, anns)
    data['tgt_out_reg'], data['tgt_out_lab'] = reg, lab
    data['tgt_overlaps'] = np.array([gen_mod.Parameters.negative_overlap, gen_mod.Parameters.positive_overlap])
    path = os.path.join(ROOT, "tests", "golden", "generator_half.npz")
    np.savez_compressed(path, **data)
    print("wrote", path, {k: getattr(v, 'shape', None) for k, v in data.items() if k.endswith('out') or 'out_' in k})


if __name__ == "__main__":
    main()
