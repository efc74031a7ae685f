"""The reference's host post-step on the detections and its on-disk formats.

  * :func:`rescale_and_cut`   ``RetinaNet.py:366-377``: ``boxes /= image_scale`` and the 0.6 score cut, as a GPU
    epilogue (``rn_rescale_cut``) on the padded ``(B, 300, .)`` outputs of ``FilterDetections`` / ``DetectionHead``;
  * :func:`read_annotations_csv`  the annotation CSV the generator consumes (``csv_generator.py:16-52``: first row
    skipped, then ``image_id, xmin, ymin, xmax, ymax, label``), grouped per image into the generator's
    ``{'bboxes', 'labels'}`` dicts (``csv_generator.py:497-512``);
  * :func:`write_detections_csv`  one row per visible detection, same column order.
"""
import csv

import numpy as np
import torch

from . import _lib


def rescale_and_cut(boxes, scores, image_scale, min_score=0.6, out=None):
    """``boxes`` (B,M,4) and ``scores`` (B,M) float32 CUDA tensors (score-sorted, padded with -1);
    ``image_scale``: one float or B floats (the resize scale of each page).  Returns ``(boxes / scale, counts)``:
    ``counts[b]`` detections of page b are at or above ``min_score`` (the reference stops at the first lower one)."""
    _lib.require_cuda()
    b = boxes if isinstance(boxes, torch.Tensor) and boxes.is_cuda else torch.as_tensor(np.asarray(boxes, np.float32)).cuda()
    s = scores if isinstance(scores, torch.Tensor) and scores.is_cuda else torch.as_tensor(np.asarray(scores, np.float32)).cuda()
    b, s = b.contiguous(), s.contiguous()
    B, M = int(s.shape[0]), int(s.shape[1])
    if tuple(b.shape) != (B, M, 4):
        raise ValueError("boxes %s does not match scores %s" % (tuple(b.shape), tuple(s.shape)))
    scale = torch.as_tensor(np.broadcast_to(np.asarray(image_scale, np.float32), (B,)).copy(), device=b.device) \
        if not isinstance(image_scale, torch.Tensor) else image_scale.to(device=b.device, dtype=torch.float32).expand(B).contiguous()
    boxes_out = torch.empty_like(b) if out is None else out
    counts = torch.empty((B,), dtype=torch.int32, device=b.device)
    _lib.check(_lib.load().rn_rescale_cut(_lib.ptr(b), _lib.ptr(s), _lib.ptr(scale), B, M, float(np.float32(min_score)),
                                          _lib.ptr(boxes_out), _lib.ptr(counts), _lib.stream_ptr(b.device)),
               "rn_rescale_cut")
    return boxes_out, counts


def read_annotations_csv(path, class_ids=None):
    """Parse the reference's annotation CSV: the first row is a header and is skipped; every other row is
    ``image_id, xmin, ymin, xmax, ymax, label``.  Returns ``{image_id: {'bboxes': (G,4) float64, 'labels': (G,)
    float64}}`` in first-appearance order.  ``class_ids`` maps label strings to ids (numeric labels pass through)."""
    pages = {}
    with open(path, newline="") as fh:
        rows = csv.reader(fh)
        next(rows, None)
        for lineno, row in enumerate(rows, start=2):
            if not row:
                continue
            if len(row) < 6:
                raise ValueError("line %d: expected 'image_id,xmin,ymin,xmax,ymax,label', got %r" % (lineno, row))
            name, label = row[0], row[5]
            try:
                box = [float(v) for v in row[1:5]]
            except ValueError:
                raise ValueError("line %d: malformed box %r" % (lineno, row[1:5]))
            if class_ids is not None and label in class_ids:
                lab = float(class_ids[label])
            else:
                try:
                    lab = float(label)
                except ValueError:
                    raise ValueError("line %d: unknown class name %r" % (lineno, label))
            entry = pages.setdefault(name, {'bboxes': [], 'labels': []})
            entry['bboxes'].append(box)
            entry['labels'].append(lab)
    return {k: {'bboxes': np.asarray(v['bboxes'], np.float64).reshape(-1, 4), 'labels': np.asarray(v['labels'], np.float64)}
            for k, v in pages.items()}


def write_detections_csv(path, image_ids, boxes, scores, labels, counts, labels_to_names=None):
    """One row per visible detection: ``image_id, xmin, ymin, xmax, ymax, label, score`` (a header row first, so the
    file reads back with :func:`read_annotations_csv`; boxes are truncated to ints like the reference's drawing code)."""
    bx = boxes.detach().cpu().numpy() if isinstance(boxes, torch.Tensor) else np.asarray(boxes)
    sc = scores.detach().cpu().numpy() if isinstance(scores, torch.Tensor) else np.asarray(scores)
    lb = labels.detach().cpu().numpy() if isinstance(labels, torch.Tensor) else np.asarray(labels)
    ct = counts.detach().cpu().numpy() if isinstance(counts, torch.Tensor) else np.asarray(counts)
    with open(path, "w", newline="") as fh:
        out = csv.writer(fh)
        out.writerow(["image_id", "xmin", "ymin", "xmax", "ymax", "label", "score"])
        for page, ident in enumerate(image_ids):
            for k in range(int(ct[page])):
                x1, y1, x2, y2 = bx[page, k].astype(int)
                lab = int(lb[page, k])
                name = labels_to_names.get(lab, lab) if labels_to_names else lab
                out.writerow([ident, x1, y1, x2, y2, name, "%.6f" % float(sc[page, k])])
