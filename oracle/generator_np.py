"""CPU restatement (numpy) of the steps either side of the hot path -- TEST INFRASTRUCTURE ONLY, never imported
by the product package.

  * ``filter_annotations``  csv_generator.py:192-218   validity rules for GT boxes
  * ``compute_inputs``      csv_generator.py:320-337   pad the page images to the batch-max shape
  * ``compute_targets``     csv_generator.py:352-370   anchors for the batch-max shape -> anchor_targets_bbox
  * ``rescale_and_cut``     RetinaNet.py:366-377       boxes /= image_scale, stop at the first score < 0.6

Pinned: ``tests/golden/make_golden_generator.py`` runs the reference's own ``Generator`` methods (imported
unmodified under stub keras/tensorflow modules, ``oracle/ref_loader.py``) and stores their outputs in
``tests/golden/generator_half.npz``; this file is checked against them bit for bit.  ``rescale_and_cut`` restates
a few lines of the entry script (not importable: it runs a model at import) and is unpinned.
"""
import warnings

import numpy as np

from . import anchors_np


def filter_annotations(image_group, annotations_group, group=None):
    """A box is dropped when x2 <= x1, y2 <= y1, x1 < 0, y1 < 0, x2 > image width or y2 > image height; every key
    of the annotation dict loses the same rows; one warning per image that had invalid boxes."""
    for i, (image, ann) in enumerate(zip(image_group, annotations_group)):
        b = ann['bboxes']
        h, w = image.shape[0], image.shape[1]
        bad = np.where((b[:, 2] <= b[:, 0]) | (b[:, 3] <= b[:, 1]) | (b[:, 0] < 0) | (b[:, 1] < 0) |
                       (b[:, 2] > w) | (b[:, 3] > h))[0]
        if len(bad):
            warnings.warn('Image with id {} (shape {}) contains the following invalid boxes: {}.'.format(
                None if group is None else group[i], image.shape, b[bad, :]))
            for k in list(ann.keys()):
                annotations_group[i][k] = np.delete(ann[k], bad, axis=0)
    return image_group, annotations_group


def compute_inputs(image_group, batch_size=None):
    """Zero batch of the max shape over the group (per axis), every image copied to the upper left corner."""
    max_shape = tuple(max(im.shape[x] for im in image_group) for x in range(3))
    n = len(image_group) if batch_size is None else batch_size
    out = np.zeros((n,) + max_shape, dtype=np.float32)
    for i, im in enumerate(image_group):
        out[i, :im.shape[0], :im.shape[1], :im.shape[2]] = im
    return out


def compute_targets(image_group, annotations_group, num_classes, anchor_params=None, shapes_callback=None,
                    negative_overlap=0.4, positive_overlap=0.5):
    max_shape = tuple(max(im.shape[x] for im in image_group) for x in range(3))
    anchors = anchors_np.anchors_for_shape(max_shape, anchor_params=anchor_params, shapes_callback=shapes_callback)
    return list(anchors_np.anchor_targets_bbox(anchors, image_group, annotations_group, num_classes,
                                               negative_overlap=negative_overlap, positive_overlap=positive_overlap))


def rescale_and_cut(boxes, scores, image_scale, min_score=0.6):
    """boxes (B,M,4) f32 / scale per page (numpy: the float32 array divided by a Python float stays float32);
    count = index of the first score < min_score in the (score-sorted) list, M if none."""
    boxes = np.array(boxes, dtype=np.float32, copy=True)
    scale = np.broadcast_to(np.asarray(image_scale, dtype=np.float64), (boxes.shape[0],))
    counts = np.zeros(boxes.shape[0], np.int32)
    for b in range(boxes.shape[0]):
        boxes[b] /= float(scale[b])
        below = np.nonzero(scores[b] < min_score)[0]
        counts[b] = below[0] if len(below) else scores.shape[1]
    return boxes, counts
