"""Drop-in for the reference's ``model/anchors.py`` -- same names, arguments and error behaviour,
the arithmetic done by kernel K1 (``csrc/anchor_targets.cu``) on a B200.

Reference seams served:
  * ``Generator(compute_anchor_targets=anchor_targets_bbox, compute_shapes=guess_shapes)``
    (``csv_generator.py:66-67``), invoked at ``csv_generator.py:361-368``;
  * ``anchors_for_shape(image_shape, anchor_params=..., shapes_callback=...)`` (``csv_generator.py:350``).

Host code here only validates, packs the ragged GT list into one pinned staging buffer and launches;
there is no CPU implementation of the path in this package.
"""

import numpy as np
import torch

from . import _lib

FLOATX = np.float32   # keras.backend.floatx() of the reference


class AnchorParameters(object):
    """model/anchors.py:7-22."""

    def __init__(self, sizes, strides, ratios, scales):
        self.sizes = sizes
        self.strides = strides
        self.ratios = ratios
        self.scales = scales

    def num_anchors(self):
        return len(self.ratios) * len(self.scales)


# model/anchors.py:28-33 -- ratios / scales are float32 on purpose (floatx)
AnchorParameters_default = AnchorParameters(
    sizes=[32, 64, 128, 256, 512],
    strides=[8, 16, 32, 64, 128],
    ratios=np.array([0.5, 1, 2], FLOATX),
    scales=np.array([2 ** 0, 2 ** (1.0 / 3.0), 2 ** (2.0 / 3.0)], FLOATX),
)


def generate_anchors(base_size=16, ratios=None, scales=None):
    """model/anchors.py:243-278: the (R*S, 4) float64 base boxes of one pyramid level.

    Host-side constant (9 boxes); uploaded once per parameter set.  ``base_size * scales`` is an fp32
    product when ``scales`` is float32, exactly as in the reference, then widened to fp64."""
    if ratios is None:
        ratios = AnchorParameters_default.ratios
    if scales is None:
        scales = AnchorParameters_default.scales
    count = len(ratios) * len(scales)
    boxes = np.zeros((count, 4))
    boxes[:, 2] = base_size * np.tile(scales, len(ratios))
    boxes[:, 3] = boxes[:, 2]
    ratio_of = np.repeat(ratios, len(scales))
    boxes[:, 2] = np.sqrt(boxes[:, 2] * boxes[:, 3] / ratio_of)
    boxes[:, 3] = boxes[:, 2] * ratio_of
    boxes[:, 0] -= boxes[:, 2] * 0.5
    boxes[:, 2] -= boxes[:, 2] * 0.5
    boxes[:, 1] -= boxes[:, 3] * 0.5
    boxes[:, 3] -= boxes[:, 3] * 0.5
    return boxes


def guess_shapes(image_shape, pyramid_levels):
    """model/anchors.py:155-165."""
    hw = np.array(image_shape[:2])
    return [(hw + 2 ** x - 1) // (2 ** x) for x in pyramid_levels]


class AnchorSpec(object):
    """What K1 needs to regenerate a set of anchors in-kernel: level table + base boxes."""

    def __init__(self, level_hw, strides, base):
        self.level_hw = np.ascontiguousarray(np.asarray(level_hw, dtype=np.int32).reshape(-1, 2))
        self.strides = np.ascontiguousarray(np.asarray(strides, dtype=np.int32).ravel())
        self.base = np.ascontiguousarray(np.asarray(base, dtype=np.float64))       # (L, A, 4)
        self.num_levels = int(self.level_hw.shape[0])
        self.per_cell = int(self.base.shape[1])
        self.num_anchors = int((self.level_hw[:, 0].astype(np.int64) * self.level_hw[:, 1]).sum() * self.per_cell)
        self._dev = {}

    def key(self):
        return (self.level_hw.tobytes(), self.strides.tobytes(), self.base.tobytes())

    def base_f64(self, device):
        t = self._dev.get(("f64", str(device)))
        if t is None:
            t = torch.from_numpy(self.base).to(device)
            self._dev[("f64", str(device))] = t
        return t

    def base_f32(self, device):
        """Base boxes cast to float32 (what ``K.variable`` does in the Anchors layer, model/layers.py:34)."""
        t = self._dev.get(("f32", str(device)))
        if t is None:
            t = torch.from_numpy(self.base.astype(np.float32)).to(device)
            self._dev[("f32", str(device))] = t
        return t


class AnchorArray(np.ndarray):
    """Read-only (N,4) float64 array returned by :func:`anchors_for_shape`.  It remembers how it was
    generated (``.spec``) so :func:`anchor_targets_bbox` can regenerate the anchors inside K1 instead of
    reading 32 bytes per anchor; any derived array (slice, copy, astype) drops the spec."""
    spec = None

    def __array_finalize__(self, obj):
        self.spec = None


def make_spec(image_shape, pyramid_levels=None, anchor_params=None, shapes_callback=None):
    if pyramid_levels is None:
        pyramid_levels = [3, 4, 5, 6, 7]
    if anchor_params is None:
        anchor_params = AnchorParameters_default
    if shapes_callback is None:
        shapes_callback = guess_shapes
    level_shapes = shapes_callback(image_shape, pyramid_levels)
    level_hw = [[int(s[0]), int(s[1])] for s in level_shapes]
    base = np.stack([generate_anchors(base_size=anchor_params.sizes[i], ratios=anchor_params.ratios,
                                      scales=anchor_params.scales) for i in range(len(pyramid_levels))], axis=0)
    return AnchorSpec(level_hw, [anchor_params.strides[i] for i in range(len(pyramid_levels))], base)


_anchor_cache = {}


def anchors_for_shape(image_shape, pyramid_levels=None, anchor_params=None, shapes_callback=None):
    """model/anchors.py:169-204.  (N,4) float64, levels P3..P7 concatenated, generated on the device
    (``rn_anchors_f64``) and cached per (level shapes, strides, base boxes)."""
    _lib.require_cuda()
    spec = make_spec(image_shape, pyramid_levels, anchor_params, shapes_callback)
    hit = _anchor_cache.get(spec.key())
    if hit is not None:
        return hit
    device = torch.device("cuda", torch.cuda.current_device())
    out = torch.empty((spec.num_anchors, 4), dtype=torch.float64, device=device)
    if spec.num_anchors:
        _, hw_p = _lib.host_ints(spec.level_hw)
        _, st_p = _lib.host_ints(spec.strides)
        _lib.check(_lib.load().rn_anchors_f64(_lib.ptr(spec.base_f64(device)), hw_p, st_p, spec.num_levels,
                                              spec.per_cell, _lib.ptr(out), _lib.stream_ptr(device)), "rn_anchors_f64")
    arr = out.cpu().numpy().view(AnchorArray)
    arr.spec = spec
    arr.flags.writeable = False
    if len(_anchor_cache) > 32:
        _anchor_cache.clear()
    _anchor_cache[spec.key()] = arr
    return arr


def shift(shape, stride, anchors):
    """model/anchors.py:208-238 for one level, on the device."""
    _lib.require_cuda()
    anchors = np.asarray(anchors, dtype=np.float64)
    spec = AnchorSpec([[int(shape[0]), int(shape[1])]], [int(stride)], anchors[None])
    device = torch.device("cuda", torch.cuda.current_device())
    out = torch.empty((spec.num_anchors, 4), dtype=torch.float64, device=device)
    if spec.num_anchors:
        _, hw_p = _lib.host_ints(spec.level_hw)
        _, st_p = _lib.host_ints(spec.strides)
        _lib.check(_lib.load().rn_anchors_f64(_lib.ptr(spec.base_f64(device)), hw_p, st_p, 1, spec.per_cell,
                                              _lib.ptr(out), _lib.stream_ptr(device)), "rn_anchors_f64")
    return out.cpu().numpy()


def _check_norm(name, value, default):
    """mean/std validation of model/anchors.py:285-298 (same messages)."""
    if value is None:
        value = np.array(default)
    if isinstance(value, (list, tuple)):
        value = np.array(value)
    elif not isinstance(value, np.ndarray):
        raise ValueError('Expected {} to be a np.ndarray, list or tuple. Received: {}'.format(name, type(value)))
    return value


# ------------------------------------------------------------------------------------------------
# GT staging: ragged annotations -> one pinned buffer -> one H2D copy
# ------------------------------------------------------------------------------------------------
class _Staging(object):
    def __init__(self):
        self.host = None
        self.event = None

    def get(self, nbytes):
        if self.event is not None:
            self.event.synchronize()          # previous async copy out of this buffer has finished
        if self.host is None or self.host.numel() < nbytes:
            self.host = torch.empty(max(nbytes, 4096), dtype=torch.uint8, pin_memory=True)
        return self.host


_staging = {}


_hw_cache = {}        # tuple of page (H, W) -> (B, 2) int32 array (pure conversion, memoised)


def pack_annotations(image_group, annotations_group, num_classes, out=None):
    """Validate like the reference (``model/anchors.py:56-60``) and flatten the ragged GT list.
    Returns ``(boxes (B,G,4) f64, labels (B,G) i32, counts (B) i32, img_hw (B,2) i32)`` numpy arrays,
    with G = max(1, max GT per page).  Labels follow numpy fancy-index rules: ``.astype(int)``
    truncation, negative values wrap over the C+1 columns, anything else raises IndexError.

    ``out``: four preallocated arrays of those dtypes, ``(B, Gcap, 4)``, ``(B, Gcap)``, ``(B,)``, ``(B, 2)`` with
    ``Gcap >= G`` (e.g. views of a pinned staging block) that are filled instead; ValueError if they do not fit."""
    assert (len(image_group) == len(annotations_group)), "The length of the images and annotations need to be equal."
    assert (len(annotations_group) > 0), "No data received to compute anchor targets for."
    for annotations in annotations_group:
        assert ('bboxes' in annotations), "Annotations should contain bboxes."
        assert ('labels' in annotations), "Annotations should contain labels."
    B = len(image_group)
    parts = [np.asarray(a['bboxes']) for a in annotations_group]
    cnt = [p.shape[0] for p in parts]
    G = max(1, max(cnt))
    if out is None:
        boxes = np.zeros((B, G, 4), dtype=np.float64)
        labels = np.zeros((B, G), dtype=np.int32)
        counts = np.empty(B, dtype=np.int32)
        img_hw = np.empty((B, 2), dtype=np.int32)
    else:
        boxes, labels, counts, img_hw = out
        if boxes.shape[0] != B or boxes.shape[1] < G:
            raise ValueError("batch of %d pages / %d GT does not fit the staging block (%d pages, %d GT)"
                             % (B, G, boxes.shape[0], boxes.shape[1]))
        boxes[...] = 0.0
        labels[...] = 0
    counts[:] = cnt
    width = num_classes + 1
    total = sum(cnt)
    if total:
        # This runs on the host once per step, so the common case -- every page's boxes a (g, 4) array and labels a
        # (g,) array -- is one concatenation, one conversion / range check and one masked assignment for the batch.
        lparts = [np.asarray(a['labels']) for a in annotations_group]
        flat = lab = None
        if all(p.ndim == 2 and p.shape[1] == 4 for p in parts) and [l.shape for l in lparts] == [(c,) for c in cnt]:
            flat = np.concatenate(parts)
            lab = np.concatenate(lparts).astype(int)
        else:                                           # ragged shapes: the reference's per-page conversions
            fl, ll = [], []
            for b in range(B):
                g = cnt[b]
                if g:
                    fl.append(np.asarray(parts[b], dtype=np.float64).reshape(g, -1)[:, :4])
                    l = lparts[b].astype(int).reshape(-1)[:g]
                    if l.size != g:
                        if l.size != 1:                 # numpy's assignment would refuse to broadcast these
                            raise ValueError("page %d: %d labels for %d boxes" % (b, l.size, g))
                        l = np.repeat(l, g)
                    ll.append(l)
            flat, lab = np.concatenate(fl), np.concatenate(ll)
        lo, hi = lab.min(), lab.max()
        if hi >= width or lo < -width:
            raise IndexError("label out of bounds for %d classes" % num_classes)
        if lo < 0:
            lab = np.where(lab < 0, lab + width, lab)
        mask = np.arange(boxes.shape[1]) < counts[:, None]       # row-major order == concatenation order
        boxes[mask] = flat
        labels[mask] = lab
    shapes = tuple(tuple(image.shape)[:2] for image in image_group)
    hw = _hw_cache.get(shapes)
    if hw is None:
        no_rule = np.iinfo(np.int32).max                # `if image.shape:` false -> no border rule
        hw = np.array([(int(s[0]), int(s[1])) if s else (no_rule, no_rule) for s in shapes], dtype=np.int32)
        if len(_hw_cache) < 64:
            _hw_cache[shapes] = hw
    img_hw[...] = hw
    return boxes, labels, counts, img_hw


def page_launch_order(boxes, out=None):
    """The order in which K1 should start the pages of a batch: heaviest first.  The kernel hands its CTAs out page by
    page, so the pages at the end of the order make the tail of the launch; a page's cost grows with the number and the
    size of its tables (``profiles/k1_page_cost.py``), and the summed table area of the packed ``(B, G, 4)`` block (padding
    rows are zero) ranks the pages well enough for that.  Returns / fills a (B,) int32 permutation; stable, deterministic."""
    area = (np.clip(boxes[:, :, 2] - boxes[:, :, 0], 0, None) * np.clip(boxes[:, :, 3] - boxes[:, :, 1], 0, None)).sum(axis=1)
    order = np.argsort(-area, kind='stable').astype(np.int32)
    if out is not None:
        out[...] = order
        return out
    return order


def upload_annotations(boxes, labels, counts, img_hw, device):
    """One pinned staging buffer, one async H2D copy; returns device views."""
    B, G = labels.shape
    nb, nl, nc, ni = boxes.nbytes, labels.nbytes, counts.nbytes, img_hw.nbytes
    total = nb + nl + nc + ni
    st = _staging.setdefault(str(device), _Staging())
    host = st.get(total)
    hv = host.numpy()
    hv[:nb] = boxes.reshape(-1).view(np.uint8)
    hv[nb:nb + nl] = labels.reshape(-1).view(np.uint8)
    hv[nb + nl:nb + nl + nc] = counts.view(np.uint8)
    hv[nb + nl + nc:total] = img_hw.reshape(-1).view(np.uint8)
    dev = torch.empty(total, dtype=torch.uint8, device=device)
    dev.copy_(host[:total], non_blocking=True)
    st.event = torch.cuda.Event()
    st.event.record(torch.cuda.current_stream(device))
    d_boxes = dev[:nb].view(torch.float64).view(B, G, 4)
    d_labels = dev[nb:nb + nl].view(torch.int32).view(B, G)
    d_counts = dev[nb + nl:nb + nl + nc].view(torch.int32)
    d_hw = dev[nb + nl + nc:total].view(torch.int32).view(B, 2)
    return d_boxes, d_labels, d_counts, d_hw


def anchor_targets_device(anchors, d_boxes, d_labels, d_counts, d_hw, num_classes,
                          negative_overlap=0.4, positive_overlap=0.5, want_argmax=False, out=None,
                          npos_total=None, npos_out=None, page_order=None, sparse_regression=False):
    """Launch K1 on GT already resident on the device.  ``anchors``: an :class:`AnchorArray` /
    :class:`AnchorSpec` (generated in-kernel) or a CUDA float64 (N,4) tensor (explicit).
    ``npos_total``: optional 1-float CUDA tensor receiving the batch's positive count (loss normaliser);
    ``npos_out``: optional (B,) int32 tensor for the per-page counts (both counters are cleared by one small kernel
    that K1 is launched behind); ``page_order``: optional (B,) int32 CUDA permutation, the order in which the
    kernel starts the pages (:func:`page_launch_order`; the results do not depend on it).
    ``sparse_regression`` (extension, ``rn_anchor_targets_sparse``): only the regression rows of state == 1 anchors are
    written -- all a smooth-L1 loss that takes the state from the label tensor ever reads (model/losses.py:72-74); the
    other rows of ``regression`` keep whatever they held.  Generated anchors, one class, no argmax.
    Returns ``(regression (B,N,5), labels (B,N,C+1), npos (B) int32, argmax (B,N) int32 | None)``."""
    lib = _lib.load()
    device = d_counts.device
    B, G = d_labels.shape
    spec = anchors if isinstance(anchors, AnchorSpec) else getattr(anchors, "spec", None)
    if spec is not None:
        N = spec.num_anchors
        explicit = None
        base = spec.base_f64(device)
        hw_keep, hw_p = _lib.host_ints(spec.level_hw)
        st_keep, st_p = _lib.host_ints(spec.strides)
        levels, per_cell = spec.num_levels, spec.per_cell
    else:
        explicit = anchors
        N = int(explicit.shape[0])
        base, hw_p, st_p, levels, per_cell = None, None, None, 0, 0
    if out is None:
        # (sparse: the rows that are not written read as zeros)
        regression = (torch.zeros if sparse_regression else torch.empty)((B, N, 5), dtype=torch.float32, device=device)
        labels = torch.empty((B, N, num_classes + 1), dtype=torch.float32, device=device)
    else:
        regression, labels = out
    npos = torch.empty((B,), dtype=torch.int32, device=device) if npos_out is None else npos_out
    argmax = torch.empty((B, N), dtype=torch.int32, device=device) if want_argmax else None
    if sparse_regression:
        if explicit is not None or num_classes != 1 or want_argmax:
            raise ValueError("sparse_regression needs generated anchors, one class and no argmax tensor")
        if N > 0:
            _lib.check(lib.rn_anchor_targets_sparse(_lib.ptr(base), hw_p, st_p, levels, per_cell, N,
                                                    _lib.ptr(d_boxes), _lib.ptr(d_labels), _lib.ptr(d_counts), _lib.ptr(d_hw),
                                                    B, G, float(np.float32(negative_overlap)), float(np.float32(positive_overlap)),
                                                    _lib.ptr(regression), _lib.ptr(labels), _lib.ptr(npos),
                                                    _lib.ptr(npos_total), _lib.ptr(page_order), _lib.stream_ptr(device)),
                       "rn_anchor_targets_sparse")
    elif N > 0:
        _lib.check(lib.rn_anchor_targets_ordered(_lib.ptr(base), hw_p, st_p, levels, per_cell,
                                                 _lib.ptr(explicit), N,
                                                 _lib.ptr(d_boxes), _lib.ptr(d_labels), _lib.ptr(d_counts), _lib.ptr(d_hw),
                                                 B, G, num_classes, float(np.float32(negative_overlap)),
                                                 float(np.float32(positive_overlap)),
                                                 _lib.ptr(regression), _lib.ptr(labels), _lib.ptr(argmax), _lib.ptr(npos),
                                                 _lib.ptr(npos_total), _lib.ptr(page_order), _lib.stream_ptr(device)),
                   "rn_anchor_targets_ordered")
    if N <= 0:
        npos.zero_()
        if npos_total is not None:
            npos_total.zero_()
    return regression, labels, npos, argmax


def _explicit_anchors(anchors, device):
    if isinstance(anchors, torch.Tensor):
        return anchors.to(device=device, dtype=torch.float64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(anchors, dtype=np.float64))).to(device)


def anchor_targets_bbox(anchors, image_group, annotations_group, num_classes,
                        negative_overlap=0.4, positive_overlap=0.5, output="numpy", return_npos=False):
    """model/anchors.py:36-92.  Returns ``(regression_batch (B,N,5), labels_batch (B,N,C+1))`` float32,
    in that order, last column = anchor state (-1 ignore / 0 background / 1 object).

    ``output="numpy"`` (default, what the reference's generator expects) copies the result to host;
    ``output="torch"`` leaves both tensors on the GPU for the loss kernel.  ``return_npos=True`` appends
    the per-page positive counts (int32 tensor / array) -- the loss normaliser K1 gets for free."""
    _lib.require_cuda()
    boxes, labels, counts, img_hw = pack_annotations(image_group, annotations_group, num_classes)
    device = torch.device("cuda", torch.cuda.current_device())
    d_boxes, d_labels, d_counts, d_hw = upload_annotations(boxes, labels, counts, img_hw, device)
    if getattr(anchors, "spec", None) is None and not isinstance(anchors, AnchorSpec):
        anchors = _explicit_anchors(anchors, device)
    regression, labels_t, npos, _ = anchor_targets_device(anchors, d_boxes, d_labels, d_counts, d_hw, num_classes,
                                                          negative_overlap, positive_overlap)
    if output == "torch":
        return (regression, labels_t, npos) if return_npos else (regression, labels_t)
    if output != "numpy":
        raise ValueError("output must be 'numpy' or 'torch'")
    h_reg = torch.empty(regression.shape, dtype=torch.float32, pin_memory=True)
    h_lab = torch.empty(labels_t.shape, dtype=torch.float32, pin_memory=True)
    h_reg.copy_(regression, non_blocking=True)
    h_lab.copy_(labels_t, non_blocking=True)
    h_npos = npos.cpu() if return_npos else None          # synchronises the stream
    if h_npos is None:
        torch.cuda.current_stream(device).synchronize()
    res = (h_reg.numpy(), h_lab.numpy())
    return res + (h_npos.numpy(),) if return_npos else res


def compute_gt_annotations(anchors, annotations, negative_overlap=0.4, positive_overlap=0.5):
    """model/anchors.py:96-117.  Returns ``(positive_indices, ignore_indices, argmax_overlaps_inds)``:
    two boolean (N,) arrays and the int64 index of the best-overlapping GT (first maximum of the fp32 IoU)."""
    _lib.require_cuda()
    annotations = np.asarray(annotations)
    device = torch.device("cuda", torch.cuda.current_device())
    ann = {'bboxes': annotations[:, :4], 'labels': np.zeros((annotations.shape[0],))}
    boxes, labels, counts, img_hw = pack_annotations([_NoShape()], [ann], 1)
    d = upload_annotations(boxes, labels, counts, img_hw, device)
    if getattr(anchors, "spec", None) is None:
        anchors = _explicit_anchors(anchors, device)
    reg, lab, _, argmax = anchor_targets_device(anchors, d[0], d[1], d[2], None, 1,
                                                negative_overlap, positive_overlap, want_argmax=True)
    state = lab[0, :, -1].cpu().numpy()
    return state == 1, state == -1, argmax[0].cpu().numpy().astype(np.int64)


class _NoShape(object):
    shape = ()


def bbox_transform(anchors, gt_boxes, mean=None, std=None):
    """model/anchors.py:282-313: row-wise corner deltas of ``gt_boxes`` w.r.t. ``anchors``, normalised by
    the anchor width/height, then ``(t - mean) / std``; float64 in and out (``rn_bbox_transform``).
    Raises the reference's ``ValueError`` for a mean/std that is not an ndarray, list or tuple."""
    mean = _check_norm('mean', mean, [0, 0, 0, 0])
    std = _check_norm('std', std, [0.2, 0.2, 0.2, 0.2])
    _lib.require_cuda()
    device = torch.device("cuda", torch.cuda.current_device())
    a = _explicit_anchors(anchors, device)
    g = _explicit_anchors(gt_boxes, device)
    if a.shape != g.shape or a.dim() != 2 or a.shape[1] != 4:
        raise ValueError("anchors and gt_boxes must both be (N, 4)")
    out = torch.empty_like(a)
    m = np.ascontiguousarray(np.broadcast_to(np.asarray(mean, dtype=np.float64), (4,)))
    s = np.ascontiguousarray(np.broadcast_to(np.asarray(std, dtype=np.float64), (4,)))
    dp = _lib.POINTER(_lib.c_double)
    _lib.check(_lib.load().rn_bbox_transform(_lib.ptr(a), _lib.ptr(g), a.shape[0], m.ctypes.data_as(dp),
                                             s.ctypes.data_as(dp), _lib.ptr(out), _lib.stream_ptr(device)),
               "rn_bbox_transform")
    return out.cpu().numpy()
