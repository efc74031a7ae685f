"""Oracle (TEST INFRASTRUCTURE): numpy restatement of the reference's anchor/target math.

Follows ``model/anchors.py`` and ``model/utils.py:180-211`` of the reference.  It is
written to reproduce the reference's *precision pipeline* exactly (which operations run in
fp32, which in fp64, and where the fp64->fp32 roundings happen), so its outputs are
bit-identical to the reference's -- this is pinned by ``tests/golden/`` fixtures generated
from the reference itself (``tests/golden/make_golden.py``).

Precision pipeline being reproduced:
  * ratios / scales are **float32** arrays (``model/anchors.py:28-33``); ``base_size*scale``
    is therefore an fp32 product that is then stored into an fp64 array (``:266``).
  * everything after that (sqrt, shifts, IoU, bbox_transform) is fp64.
  * the IoU matrix is *stored* as float32 (``model/utils.py:190,209``); argmax and both
    threshold comparisons run on those fp32 values, the thresholds being weak Python floats
    (so they are compared as ``np.float32(0.5)`` / ``np.float32(0.4)``).
"""
import numpy as np

FLOATX = np.float32  # keras.backend.floatx() of the reference


class AnchorParameters(object):
    """model/anchors.py:7-22."""

    def __init__(self, sizes, strides, ratios, scales):
        self.sizes, self.strides, self.ratios, self.scales = sizes, strides, ratios, scales

    def num_anchors(self):
        return len(self.ratios) * len(self.scales)


# model/anchors.py:28-33
AnchorParameters_default = AnchorParameters(
    sizes=[32, 64, 128, 256, 512],
    strides=[8, 16, 32, 64, 128],
    ratios=np.array([0.5, 1, 2], FLOATX),
    scales=np.array([2 ** 0, 2 ** (1.0 / 3.0), 2 ** (2.0 / 3.0)], FLOATX),
)


def generate_anchors(base_size=16, ratios=None, scales=None):
    """model/anchors.py:243-278.  (R*S, 4) float64 boxes centred on the origin, ratio-major."""
    ratios = AnchorParameters_default.ratios if ratios is None else ratios
    scales = AnchorParameters_default.scales if scales is None else scales
    n_r, n_s = len(ratios), len(scales)
    per_anchor_scale = np.tile(scales, n_r)          # scale-minor
    per_anchor_ratio = np.repeat(ratios, n_s)        # ratio-major
    side = np.zeros(n_r * n_s)                       # fp64 store of an fp32 (or promoted) product
    side[:] = base_size * per_anchor_scale
    area = side * side
    w = np.sqrt(area / per_anchor_ratio)
    h = w * per_anchor_ratio
    out = np.zeros((n_r * n_s, 4))
    out[:, 0] = 0.0 - w * 0.5
    out[:, 1] = 0.0 - h * 0.5
    out[:, 2] = w - w * 0.5
    out[:, 3] = h - h * 0.5
    return out


def guess_shapes(image_shape, pyramid_levels):
    """model/anchors.py:155-165: ceil-divide the (H, W) by 2**level."""
    hw = np.array(image_shape[:2])
    return [(hw + 2 ** lvl - 1) // (2 ** lvl) for lvl in pyramid_levels]


def shift(shape, stride, anchors):
    """model/anchors.py:208-238.  Cell-major, anchor-minor: row = (y*W + x)*A + a."""
    cx = (np.arange(0, shape[1]) + 0.5) * stride
    cy = (np.arange(0, shape[0]) + 0.5) * stride
    gx, gy = np.meshgrid(cx, cy)
    centres = np.stack([gx.ravel(), gy.ravel(), gx.ravel(), gy.ravel()], axis=1)   # (K, 4)
    out = anchors[None, :, :] + centres[:, None, :]                                  # (K, A, 4)
    return out.reshape((-1, 4))


def anchors_for_shape(image_shape, pyramid_levels=None, anchor_params=None, shapes_callback=None):
    """model/anchors.py:169-204.  Levels concatenated P3 -> P7."""
    pyramid_levels = [3, 4, 5, 6, 7] if pyramid_levels is None else pyramid_levels
    anchor_params = AnchorParameters_default if anchor_params is None else anchor_params
    shapes_callback = guess_shapes if shapes_callback is None else shapes_callback
    level_shapes = shapes_callback(image_shape, pyramid_levels)
    per_level = [np.zeros((0, 4))]
    for i, _ in enumerate(pyramid_levels):
        base = generate_anchors(base_size=anchor_params.sizes[i], ratios=anchor_params.ratios,
                                scales=anchor_params.scales)
        per_level.append(shift(level_shapes[i], anchor_params.strides[i], base))
    return np.concatenate(per_level, axis=0)


def compute_overlap(boxes1, boxes2):
    """model/utils.py:180-211.  IoU without the +1 pixel convention; the arithmetic runs in the
    input dtype (fp64 on the target path) and the result is rounded into a **float32** matrix."""
    out = np.zeros((len(boxes1), len(boxes2)), dtype=np.float32)
    if len(boxes1) == 0 or len(boxes2) == 0:
        return out
    area1 = (boxes1[:, 2] - boxes1[:, 0]) * (boxes1[:, 3] - boxes1[:, 1])
    area2 = (boxes2[:, 2] - boxes2[:, 0]) * (boxes2[:, 3] - boxes2[:, 1])
    for j in range(len(boxes2)):
        g = boxes2[j]
        iw = np.maximum(0, np.minimum(boxes1[:, 2], g[2]) - np.maximum(boxes1[:, 0], g[0]))
        ih = np.maximum(0, np.minimum(boxes1[:, 3], g[3]) - np.maximum(boxes1[:, 1], g[1]))
        inter = iw * ih
        out[:, j] = inter / (area1 + area2[j] - inter)
    return out


def compute_gt_annotations(anchors, annotations, negative_overlap=0.4, positive_overlap=0.5):
    """model/anchors.py:96-117.  argmax is first-max on the fp32 IoU matrix; ``positive`` is
    ``>=`` and ``ignore`` is strict ``>`` (both evaluated in fp32)."""
    iou = compute_overlap(anchors.astype(np.float64), annotations.astype(np.float64))
    best = np.argmax(iou, axis=1)
    best_iou = iou[np.arange(iou.shape[0]), best]
    positive = best_iou >= positive_overlap
    ignore = (best_iou > negative_overlap) & ~positive
    return positive, ignore, best


def bbox_transform(anchors, gt_boxes, mean=None, std=None):
    """model/anchors.py:282-313.  Corner deltas normalised by anchor width/height, then
    ``(t - mean) / std``; all in the input dtype (fp64 on the target path)."""
    mean = np.array([0, 0, 0, 0]) if mean is None else mean
    std = np.array([0.2, 0.2, 0.2, 0.2]) if std is None else std
    if isinstance(mean, (list, tuple)):
        mean = np.array(mean)
    elif not isinstance(mean, np.ndarray):
        raise ValueError('Expected mean to be a np.ndarray, list or tuple. Received: {}'.format(type(mean)))
    if isinstance(std, (list, tuple)):
        std = np.array(std)
    elif not isinstance(std, np.ndarray):
        raise ValueError('Expected std to be a np.ndarray, list or tuple. Received: {}'.format(type(std)))
    aw = anchors[:, 2] - anchors[:, 0]
    ah = anchors[:, 3] - anchors[:, 1]
    cols = [(gt_boxes[:, 0] - anchors[:, 0]) / aw,
            (gt_boxes[:, 1] - anchors[:, 1]) / ah,
            (gt_boxes[:, 2] - anchors[:, 2]) / aw,
            (gt_boxes[:, 3] - anchors[:, 3]) / ah]
    return (np.stack(cols, axis=1) - mean) / std


def anchor_targets_bbox(anchors, image_group, annotations_group, num_classes,
                        negative_overlap=0.4, positive_overlap=0.5):
    """model/anchors.py:36-92.  Returns ``(regression_batch (B,N,5), labels_batch (B,N,C+1))``,
    both float32, last column = anchor state (-1 ignore / 0 background / 1 object).

    Order of effects per image (it matters): states from IoU -> one-hot class at positives ->
    regression targets for *every* anchor against its argmax GT -> anchors whose centre lies
    at/after the image's own (pre-padding) width/height forced to state -1 in both outputs
    (``:85-90``; the one-hot of an overridden positive stays set)."""
    assert len(image_group) == len(annotations_group), "The length of the images and annotations need to be equal."
    assert len(annotations_group) > 0, "No data received to compute anchor targets for."
    for ann in annotations_group:
        assert 'bboxes' in ann, "Annotations should contain bboxes."
        assert 'labels' in ann, "Annotations should contain labels."
    n_img, n_anchor = len(image_group), anchors.shape[0]
    regression = np.zeros((n_img, n_anchor, 5), dtype=FLOATX)
    labels = np.zeros((n_img, n_anchor, num_classes + 1), dtype=FLOATX)
    centre_x = (anchors[:, 0] + anchors[:, 2]) / 2
    centre_y = (anchors[:, 1] + anchors[:, 3]) / 2
    for b, (image, ann) in enumerate(zip(image_group, annotations_group)):
        if ann['bboxes'].shape[0]:
            pos, ign, best = compute_gt_annotations(anchors, ann['bboxes'], negative_overlap, positive_overlap)
            labels[b, ign, -1] = -1
            labels[b, pos, -1] = 1
            regression[b, ign, -1] = -1
            regression[b, pos, -1] = 1
            labels[b, pos, ann['labels'][best[pos]].astype(int)] = 1
            regression[b, :, :-1] = bbox_transform(anchors, ann['bboxes'][best, :])
        if image.shape:
            outside = np.logical_or(centre_x >= image.shape[1], centre_y >= image.shape[0])
            labels[b, outside, -1] = -1
            regression[b, outside, -1] = -1
    return regression, labels
