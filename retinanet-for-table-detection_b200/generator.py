"""The batching steps either side of kernel K1, with the reference generator's names and behaviour
(``csv_generator.py``):

  * :func:`filter_annotations`  ``Generator.filter_annotations`` (:192-218): GT validity rules (host, numpy);
  * :func:`compute_inputs`      ``Generator.compute_inputs`` (:320-337): pad pages to the batch-max shape (host);
  * :func:`compute_targets`     ``Generator.compute_targets`` (:352-370): anchors for the batch-max shape, then
    ``anchor_targets_bbox`` -- on the GPU (K1); the anchors are never materialised.

The image tensors themselves belong to the backbone's input pipeline and stay on the host here.
"""
import warnings

import numpy as np

from . import anchors as _anchors


def filter_annotations(image_group, annotations_group, group=None):
    """Drop GT boxes with ``x2 <= x1``, ``y2 <= y1``, ``x1 < 0``, ``y1 < 0``, ``x2 > width`` or ``y2 > height``
    (every key of the annotation dict loses the same rows), warning once per affected image.  In place, like the
    reference; returns ``(image_group, annotations_group)``."""
    ids = range(len(image_group)) if group is None else group
    for slot, (image, ann, ident) in enumerate(zip(image_group, annotations_group, ids)):
        boxes = np.asarray(ann['bboxes'])
        if boxes.size == 0:
            continue
        height, width = image.shape[0], image.shape[1]
        x1, y1, x2, y2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
        invalid = (x2 <= x1) | (y2 <= y1) | (x1 < 0) | (y1 < 0) | (x2 > width) | (y2 > height)
        if not invalid.any():
            continue
        warnings.warn('Image with id {} (shape {}) contains the following invalid boxes: {}.'.format(
            ident, image.shape, boxes[invalid, :]))
        keep = ~invalid
        for key in list(ann.keys()):
            annotations_group[slot][key] = np.asarray(ann[key])[keep]
    return image_group, annotations_group


def compute_inputs(image_group, batch_size=None, dtype=np.float32):
    """Zero-filled ``(batch, Hmax, Wmax, Cmax)`` array (``keras.backend.floatx()`` = float32) with every image in
    its upper-left corner.  ``batch_size`` defaults to the group size (the reference uses its configured batch
    size, which equals the group size)."""
    dims = [max(int(im.shape[axis]) for im in image_group) for axis in range(3)]
    count = len(image_group) if batch_size is None else int(batch_size)
    batch = np.zeros([count] + dims, dtype=dtype)
    for slot, im in enumerate(image_group):
        batch[slot, :im.shape[0], :im.shape[1], :im.shape[2]] = im
    return batch


def compute_targets(image_group, annotations_group, num_classes, anchor_params=None, shapes_callback=None,
                    negative_overlap=0.4, positive_overlap=0.5, output="numpy"):
    """``[regression_batch (B,N,5), labels_batch (B,N,C+1)]`` for the anchors of the batch-max page shape; each
    page's border-ignore rule uses its own shape (model/anchors.py:85-90).  ``output="torch"`` keeps the two
    tensors on the GPU."""
    max_shape = tuple(max(int(im.shape[axis]) for im in image_group) for axis in range(3))
    anchors = _anchors.anchors_for_shape(max_shape, anchor_params=anchor_params, shapes_callback=shapes_callback)
    regression, labels = _anchors.anchor_targets_bbox(anchors, image_group, annotations_group, num_classes,
                                                      negative_overlap=negative_overlap,
                                                      positive_overlap=positive_overlap, output=output)
    return [regression, labels]
