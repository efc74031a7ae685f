"""Oracle (TEST INFRASTRUCTURE): numpy fp32 restatement of the reference's detection-head layers.

Follows ``model/utils.py:51-80`` (TF ``shift``), ``:84-112`` (``bbox_transform_inv``) and
``model/layers.py`` (``Anchors`` :42-53, ``RegressBoxes`` :136-138, ``ClipBoxes`` :157-171,
``filter_detections`` :177-264, ``FilterDetections.call`` :298-332).

PARITY UNPINNED: these run as TensorFlow graph ops in the reference; TensorFlow is not
installed here and the reference has no golden vectors.  Third-party semantics restated
(and cross-checked offline against ``torchvision.ops.nms`` in ``tests/``):

``tf.image.non_max_suppression(boxes, scores, max_output_size, iou_threshold)``
(``model/layers.py:211``; TF >= 1.14 ``non_max_suppression_op.cc``):
  * greedy; candidates are visited by (score descending, index ascending);
  * a candidate is dropped iff its IoU with *any already selected* box is ``> iou_threshold``
    (strict), otherwise it is selected; stops once ``max_output_size`` boxes are selected;
    returns indices in selection order;
  * IoU in float32: each box's corners are normalised with min/max per axis;
    ``area = (ymax - ymin) * (xmax - xmin)``; **if either area <= 0 the IoU is 0**;
    ``inter = max(dy, 0) * max(dx, 0)``; ``iou = inter / (area_i + area_j - inter)``.
``tf.nn.top_k`` (``model/layers.py:241``): descending, ties -> lower index first (stable).
``tf.where(greater(scores, thr))`` (``:202``): strict ``>``, ascending index order.
``tf.clip_by_value(x, lo, hi)`` (``:166-169``): ``min(max(x, lo), hi)``.
"""
import numpy as np

from . import anchors_np

F32 = np.float32


def shift_f32(shape, stride, anchors):
    """model/utils.py:51-80.  float32 centres ``(arange + 0.5) * stride`` added to float32 base
    anchors; cell-major, anchor-minor."""
    anchors = np.asarray(anchors, dtype=F32)
    cx = (np.arange(0, shape[1], dtype=F32) + F32(0.5)) * F32(stride)
    cy = (np.arange(0, shape[0], dtype=F32) + F32(0.5)) * F32(stride)
    gx, gy = np.meshgrid(cx, cy)
    centres = np.stack([gx.ravel(), gy.ravel(), gx.ravel(), gy.ravel()], axis=1)
    return (anchors[None, :, :] + centres[:, None, :]).reshape((-1, 4)).astype(F32)


def anchors_layer(size, stride, ratios, scales, feature_hw, batch):
    """model/layers.py:7-53.  Base anchors are computed in fp64 and cast to float32
    (``K.variable``, :34) before shifting; output tiled to (batch, H*W*A, 4)."""
    base = anchors_np.generate_anchors(base_size=size, ratios=ratios, scales=scales).astype(F32)
    one = shift_f32(feature_hw, stride, base)
    return np.tile(one[None], (batch, 1, 1))


def all_anchors_f32(image_hw, batch=1, pyramid_levels=None, anchor_params=None):
    """model/defineModel.py:271-293: per-level Anchors layers concatenated P3..P7 (feature
    shapes follow ``guess_shapes``, which is what the ResNet/FPN strides produce)."""
    pyramid_levels = [3, 4, 5, 6, 7] if pyramid_levels is None else pyramid_levels
    ap = anchors_np.AnchorParameters_default if anchor_params is None else anchor_params
    shapes = anchors_np.guess_shapes(image_hw, pyramid_levels)
    per = [anchors_layer(ap.sizes[i], ap.strides[i], ap.ratios, ap.scales, shapes[i], batch)
           for i in range(len(pyramid_levels))]
    return np.concatenate(per, axis=1)


def bbox_transform_inv(boxes, deltas, mean=None, std=None):
    """model/utils.py:84-112, float32, evaluation order as written:
    ``x1 = b_x1 + (d0 * std0 + mean0) * width`` etc."""
    mean = [0, 0, 0, 0] if mean is None else mean
    std = [0.2, 0.2, 0.2, 0.2] if std is None else std
    boxes = np.asarray(boxes, dtype=F32)
    deltas = np.asarray(deltas, dtype=F32)
    w = boxes[:, :, 2] - boxes[:, :, 0]
    h = boxes[:, :, 3] - boxes[:, :, 1]
    x1 = boxes[:, :, 0] + (deltas[:, :, 0] * F32(std[0]) + F32(mean[0])) * w
    y1 = boxes[:, :, 1] + (deltas[:, :, 1] * F32(std[1]) + F32(mean[1])) * h
    x2 = boxes[:, :, 2] + (deltas[:, :, 2] * F32(std[2]) + F32(mean[2])) * w
    y2 = boxes[:, :, 3] + (deltas[:, :, 3] * F32(std[3]) + F32(mean[3])) * h
    return np.stack([x1, y1, x2, y2], axis=2).astype(F32)


def clip_boxes(image_hw, boxes):
    """model/layers.py:157-171: clip x to [0, W], y to [0, H] (the padded input tensor's shape)."""
    boxes = np.asarray(boxes, dtype=F32)
    H, W = F32(image_hw[0]), F32(image_hw[1])
    z = F32(0)
    x1 = np.minimum(np.maximum(boxes[:, :, 0], z), W)
    y1 = np.minimum(np.maximum(boxes[:, :, 1], z), H)
    x2 = np.minimum(np.maximum(boxes[:, :, 2], z), W)
    y2 = np.minimum(np.maximum(boxes[:, :, 3], z), H)
    return np.stack([x1, y1, x2, y2], axis=2)


def _iou_one_vs_many(box, area, boxes, areas):
    lo0 = np.maximum(box[0], boxes[:, 0])
    lo1 = np.maximum(box[1], boxes[:, 1])
    hi0 = np.minimum(box[2], boxes[:, 2])
    hi1 = np.minimum(box[3], boxes[:, 3])
    inter = np.maximum(hi0 - lo0, F32(0)) * np.maximum(hi1 - lo1, F32(0))
    with np.errstate(divide='ignore', invalid='ignore'):
        iou = inter / (area + areas - inter)
    return np.where((area <= 0) | (areas <= 0), F32(0), iou)


def non_max_suppression(boxes, scores, max_output_size, iou_threshold):
    """Restatement of ``tf.image.non_max_suppression`` (see module docstring).  Returns int64
    indices into ``boxes`` in selection order."""
    boxes = np.asarray(boxes, dtype=F32).reshape(-1, 4)
    scores = np.asarray(scores, dtype=F32).reshape(-1)
    n = boxes.shape[0]
    if n == 0 or max_output_size <= 0:
        return np.zeros((0,), dtype=np.int64)
    # corner normalisation (first/third coordinate pair is one axis, second/fourth the other)
    c = np.stack([np.minimum(boxes[:, 0], boxes[:, 2]), np.minimum(boxes[:, 1], boxes[:, 3]),
                  np.maximum(boxes[:, 0], boxes[:, 2]), np.maximum(boxes[:, 1], boxes[:, 3])], axis=1)
    areas = (c[:, 2] - c[:, 0]) * (c[:, 3] - c[:, 1])
    order = np.argsort(-scores, kind='stable')          # score desc, index asc
    cs = c[order]
    ca = areas[order]
    dead = np.zeros(n, dtype=bool)
    thr = F32(iou_threshold)
    picked = []
    for r in range(n):
        if dead[r]:
            continue
        picked.append(order[r])
        if len(picked) >= max_output_size:
            break
        if r + 1 < n:
            iou = _iou_one_vs_many(cs[r], ca[r], cs[r + 1:], ca[r + 1:])
            dead[r + 1:] |= iou > thr
    return np.asarray(picked, dtype=np.int64)


def filter_detections(boxes, classification, other=None, class_specific_filter=True, nms=True,
                      score_threshold=0.05, max_detections=300, nms_threshold=0.5):
    """model/layers.py:177-264 for ONE image.  Returns ``[boxes (M,4) f32, scores (M,) f32,
    labels (M,) i32, *other]`` padded with -1 to ``M = max_detections``; additionally the
    selected anchor indices (int64, -1 padded) are returned as the last element so tests can
    check them bit-for-bit."""
    other = [] if other is None else other
    boxes = np.asarray(boxes, dtype=F32)
    classification = np.asarray(classification, dtype=F32)
    thr = F32(score_threshold)

    def _one(scores, labels):
        idx = np.nonzero(scores > thr)[0]                                        # :202
        if nms:
            keep = non_max_suppression(boxes[idx], scores[idx], max_detections, nms_threshold)  # :211
            idx = idx[keep]
        return np.stack([idx, labels[idx]], axis=1)                              # :217-218

    if class_specific_filter:
        parts = []
        for c in range(classification.shape[1]):                                 # :226-229
            parts.append(_one(classification[:, c], np.full((classification.shape[0],), c, dtype=np.int64)))
        pairs = np.concatenate(parts, axis=0) if parts else np.zeros((0, 2), dtype=np.int64)
    else:
        pairs = _one(classification.max(axis=1), classification.argmax(axis=1).astype(np.int64))  # :234-236

    sel_scores = classification[pairs[:, 0], pairs[:, 1]]                        # :239
    k = min(max_detections, sel_scores.shape[0])
    top = np.argsort(-sel_scores, kind='stable')[:k]                             # :241
    sel_scores = sel_scores[top]
    anchor_idx = pairs[top, 0]
    labels = pairs[top, 1]
    out_boxes = boxes[anchor_idx]
    out_other = [np.asarray(o)[anchor_idx] for o in other]
    pad = max(0, max_detections - k)                                             # :250-255
    out_boxes = np.concatenate([out_boxes, np.full((pad, 4), -1, dtype=F32)], axis=0)
    out_scores = np.concatenate([sel_scores, np.full((pad,), -1, dtype=F32)])
    out_labels = np.concatenate([labels, np.full((pad,), -1, dtype=np.int64)]).astype(np.int32)
    out_other = [np.concatenate([o, np.full((pad,) + o.shape[1:], -1, dtype=o.dtype)], axis=0) for o in out_other]
    out_idx = np.concatenate([anchor_idx, np.full((pad,), -1, dtype=np.int64)])
    return [out_boxes, out_scores, out_labels] + out_other + [out_idx]


def filter_detections_batch(boxes, classification, other=None, **kw):
    """model/layers.py:298-332: ``filter_detections`` mapped over the batch and stacked."""
    other = [] if other is None else other
    per_image = [filter_detections(boxes[b], classification[b], [o[b] for o in other], **kw)
                 for b in range(boxes.shape[0])]
    return [np.stack([r[i] for r in per_image], axis=0) for i in range(len(per_image[0]))]


def detect(image_hw, regression, classification, anchor_params=None, **kw):
    """The inference wiring of ``retinanet_bbox`` (model/defineModel.py:329-350):
    anchors -> RegressBoxes -> ClipBoxes -> FilterDetections."""
    B = regression.shape[0]
    anchors = all_anchors_f32(image_hw, batch=B, anchor_params=anchor_params)
    boxes = clip_boxes(image_hw, bbox_transform_inv(anchors, regression))
    return filter_detections_batch(boxes, classification, **kw)
