"""Drop-in for the hot-path layers of the reference's ``model/layers.py``: ``Anchors``, ``RegressBoxes``,
``ClipBoxes``, ``FilterDetections`` and the function ``filter_detections`` -- same constructor keywords,
list-style inputs, output order, ``get_config`` / ``compute_output_shape`` -- as ``torch.nn.Module``s whose
``forward`` (alias ``call``) launches the CUDA kernels of ``csrc/detect_layers.cu`` and
``csrc/filter_detections.cu``.  ``UpsampleLike`` (FPN) is outside the path.

``custom_objects`` mirrors the dictionary ``load_model`` is given in the reference
(``model/defineModel.py:15-24``).  :class:`DetectionHead` is the wiring of ``retinanet_bbox``
(``model/defineModel.py:329-350``) on the fused kernel path (anchors generated in-kernel, decode + clip +
threshold in one pass).

Tensors are channels-last like Keras' default: a feature map / image is (B, H, W, C).
"""
import numpy as np
import torch

from . import _lib
from . import anchors as _anchors


def _cuda(x, dtype=torch.float32):
    _lib.require_cuda()
    if isinstance(x, torch.Tensor):
        if not x.is_cuda:
            x = x.cuda()
        return x.to(dtype).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x))).cuda().to(dtype).contiguous()


def _shape_of(x):
    return tuple(x.shape) if hasattr(x, "shape") else tuple(x)


class _Layer(torch.nn.Module):
    def __init__(self, name=None, **kwargs):
        super().__init__()
        self.name = name

    def call(self, inputs, **kwargs):
        return self.forward(inputs, **kwargs)

    def get_config(self):
        return {'name': self.name} if self.name is not None else {}


class Anchors(_Layer):
    """model/layers.py:7-75.  ``Anchors(size, stride, ratios, scales)(features)`` -> (B, H*W*A, 4) float32.
    As in the reference, ``ratios`` / ``scales`` must be given (its ``None`` defaults are broken:
    they reference a non-existent ``AnchorParameters.default``); here ``None`` falls back to
    ``AnchorParameters_default`` instead of raising AttributeError."""

    def __init__(self, size, stride, ratios=None, scales=None, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.size = size
        self.stride = stride
        self.ratios = _anchors.AnchorParameters_default.ratios if ratios is None else np.array(ratios) if isinstance(ratios, list) else ratios
        self.scales = _anchors.AnchorParameters_default.scales if scales is None else np.array(scales) if isinstance(scales, list) else scales
        self.num_anchors = len(self.ratios) * len(self.scales)
        # generate_anchors in fp64, cast to floatx like K.variable (model/layers.py:34)
        self.base = _anchors.generate_anchors(base_size=size, ratios=self.ratios, scales=self.scales)

    def forward(self, inputs, **kwargs):
        shape = _shape_of(inputs)            # (B, H, W, C)
        B, H, W = int(shape[0]), int(shape[1]), int(shape[2])
        spec = _anchors.AnchorSpec([[H, W]], [int(self.stride)], self.base[None])
        device = inputs.device if isinstance(inputs, torch.Tensor) and inputs.is_cuda else torch.device("cuda", torch.cuda.current_device())
        _lib.require_cuda()
        out = torch.empty((B, spec.num_anchors, 4), dtype=torch.float32, device=device)
        if out.numel():
            hw_keep, hw_p = _lib.host_ints(spec.level_hw)
            st_keep, st_p = _lib.host_ints(spec.strides)
            _lib.check(_lib.load().rn_anchors_f32(_lib.ptr(spec.base_f32(device)), hw_p, st_p, 1, spec.per_cell, B,
                                                  _lib.ptr(out), _lib.stream_ptr(device)), "rn_anchors_f32")
        return out

    def compute_output_shape(self, input_shape):
        if None not in input_shape[1:]:
            return (input_shape[0], int(np.prod(input_shape[1:3])) * self.num_anchors, 4)
        return (input_shape[0], None, 4)

    def get_config(self):
        config = super().get_config()
        config.update({'size': self.size, 'stride': self.stride,
                       'ratios': np.asarray(self.ratios).tolist(), 'scales': np.asarray(self.scales).tolist()})
        return config


def _norm_arg(name, value, default):
    if value is None:
        value = np.array(default)
    if isinstance(value, (list, tuple)):
        value = np.array(value)
    elif not isinstance(value, np.ndarray):
        raise ValueError('Expected {} to be a np.ndarray, list or tuple. Received: {}'.format(name, type(value)))
    return value


class RegressBoxes(_Layer):
    """model/layers.py:107-150: ``RegressBoxes(mean, std)([anchors, regression])`` -> boxes (B,N,4)."""

    def __init__(self, mean=None, std=None, *args, **kwargs):
        self.mean = _norm_arg('mean', mean, [0, 0, 0, 0])
        self.std = _norm_arg('std', std, [0.2, 0.2, 0.2, 0.2])
        super().__init__(*args, **kwargs)

    def forward(self, inputs, **kwargs):
        anchors, regression = inputs
        a, d = _cuda(anchors), _cuda(regression)
        if a.shape != d.shape or a.shape[-1] != 4:
            raise ValueError("anchors %s and regression %s must both be (B, N, 4)" % (tuple(a.shape), tuple(d.shape)))
        out = torch.empty_like(a)
        m_keep, mp = _lib.host_floats(self.mean)
        s_keep, sp = _lib.host_floats(self.std)
        _lib.check(_lib.load().rn_regress_boxes(_lib.ptr(a), _lib.ptr(d), a.numel() // 4, mp, sp, _lib.ptr(out),
                                                _lib.stream_ptr(a.device)), "rn_regress_boxes")
        return out

    def compute_output_shape(self, input_shape):
        return input_shape[0]

    def get_config(self):
        config = super().get_config()
        config.update({'mean': self.mean.tolist(), 'std': self.std.tolist()})
        return config


class ClipBoxes(_Layer):
    """model/layers.py:153-174: ``ClipBoxes()([image, boxes])``; x clipped to [0, W], y to [0, H] where
    (H, W) is the shape of the (padded, channels-last) image tensor."""

    def forward(self, inputs, **kwargs):
        image, boxes = inputs
        shape = _shape_of(image)
        H, W = float(shape[1]), float(shape[2])
        b = _cuda(boxes)
        out = torch.empty_like(b)
        _lib.check(_lib.load().rn_clip_boxes(_lib.ptr(b), b.numel() // 4, W, H, _lib.ptr(out),
                                             _lib.stream_ptr(b.device)), "rn_clip_boxes")
        return out

    def compute_output_shape(self, input_shape):
        return input_shape[1]


def _gather_other(other, indices, max_detections):
    """``other`` tensors follow the selected anchors (model/layers.py:247,255): gather + pad with -1 (``rn_gather_other``;
    float32 like the reference's ``keras.backend.floatx()`` tensors)."""
    outs = []
    if not other:
        return outs
    lib = _lib.load()
    B, M = int(indices.shape[0]), int(indices.shape[1])
    for o in other:
        o = o if isinstance(o, torch.Tensor) else torch.as_tensor(np.asarray(o))
        o = o.to(device=indices.device, dtype=torch.float32).contiguous()
        if o.dim() < 2 or int(o.shape[0]) != B:
            raise ValueError("other tensors must be (B, N, ...); got %s" % (tuple(o.shape),))
        N = int(o.shape[1])
        D = int(np.prod(o.shape[2:])) if o.dim() > 2 else 1
        out = torch.empty((B, M) + tuple(o.shape[2:]), dtype=torch.float32, device=indices.device)
        _lib.check(lib.rn_gather_other(_lib.ptr(o), _lib.ptr(indices), B, N, M, D, _lib.ptr(out), _lib.stream_ptr(indices.device)),
                   "rn_gather_other")
        outs.append(out)
    return outs


def filter_outputs(B, M, device):
    """The five result tensors of one filter call: boxes (B,M,4) f32, scores (B,M) f32, labels (B,M) i32, selected anchor
    indices (B,M) i32, per-page status (B) i32."""
    return (torch.empty((B, M, 4), dtype=torch.float32, device=device), torch.empty((B, M), dtype=torch.float32, device=device),
            torch.empty((B, M), dtype=torch.int32, device=device), torch.empty((B, M), dtype=torch.int32, device=device),
            torch.empty((B,), dtype=torch.int32, device=device))


def _run_filter(boxes, classification, class_specific_filter, nms, score_threshold, max_detections,
                nms_threshold, pre_nms_top_k, cand_cap, decode=None, out=None, workspace=None):
    """Shared launcher.  ``decode`` = None (boxes given) or dict(spec, regression, mean, std, clip_hw).
    ``out`` (see :func:`filter_outputs`) / ``workspace``: caller-owned static buffers (``pipeline.DetectionStep``);
    allocated per call otherwise."""
    lib = _lib.load()
    cls = _cuda(classification)
    device = cls.device
    B, N, C = int(cls.shape[0]), int(cls.shape[1]), int(cls.shape[2])
    M = int(max_detections)
    cap = N if cand_cap is None else int(cand_cap)
    out_boxes, out_scores, out_labels, out_idx, status = filter_outputs(B, M, device) if out is None else out
    ws_bytes = int(lib.rn_filter_workspace_bytes(B, N, C, int(bool(class_specific_filter)), cap, M))
    ws = _lib.scratch("filter", ws_bytes, device) if workspace is None else workspace
    if ws.numel() < ws_bytes:
        raise ValueError("filter workspace too small: %d < %d bytes" % (ws.numel(), ws_bytes))
    thr = float(np.float32(score_threshold))
    nthr = float(np.float32(nms_threshold))
    if decode is None:
        bx = _cuda(boxes)
        if tuple(bx.shape) != (B, N, 4):
            raise ValueError("boxes %s does not match classification %s" % (tuple(bx.shape), tuple(cls.shape)))
        _lib.check(lib.rn_filter_detections(_lib.ptr(bx), _lib.ptr(cls), B, N, C, int(bool(class_specific_filter)),
                                            int(bool(nms)), thr, nthr, M, int(pre_nms_top_k), cap,
                                            _lib.ptr(out_boxes), _lib.ptr(out_scores), _lib.ptr(out_labels),
                                            _lib.ptr(out_idx), _lib.ptr(status), _lib.ptr(ws), ws_bytes,
                                            _lib.stream_ptr(device)), "rn_filter_detections")
    else:
        spec = decode['spec']
        reg = decode['regression']
        # a PINNED host tensor is used in place: K3 reads the 16-byte regression row of a candidate only (~2.5 % of
        # the anchors), so fetching those rows over PCIe (unified addressing) beats copying the whole (B,N,4) tensor
        in_place = isinstance(reg, torch.Tensor) and not reg.is_cuda and reg.is_pinned() and reg.dtype == torch.float32 \
            and reg.is_contiguous()
        if not in_place:
            reg = _cuda(reg)
        if tuple(reg.shape) != (B, N, 4) or spec.num_anchors != N:
            raise ValueError("regression %s / anchors (%d) do not match classification %s"
                             % (tuple(reg.shape), spec.num_anchors, tuple(cls.shape)))
        hw_keep, hw_p = _lib.host_ints(spec.level_hw)
        st_keep, st_p = _lib.host_ints(spec.strides)
        m_keep, mp = _lib.host_floats(decode['mean'])
        s_keep, sp = _lib.host_floats(decode['std'])
        H, W = decode['clip_hw']
        _lib.check(lib.rn_decode_filter_detections(_lib.ptr(spec.base_f32(device)), hw_p, st_p, spec.num_levels,
                                                   spec.per_cell, _lib.ptr(reg), _lib.ptr(cls), B, N, C, mp, sp,
                                                   float(W), float(H), int(bool(class_specific_filter)),
                                                   int(bool(nms)), thr, nthr, M, int(pre_nms_top_k), cap,
                                                   _lib.ptr(out_boxes), _lib.ptr(out_scores), _lib.ptr(out_labels),
                                                   _lib.ptr(out_idx), _lib.ptr(status), _lib.ptr(ws), ws_bytes,
                                                   _lib.stream_ptr(device)), "rn_decode_filter_detections")
    return out_boxes, out_scores, out_labels, out_idx, status


def _raise_on_overflow(status, cand_cap):
    if cand_cap is not None and bool((status != 0).any().item()):
        raise _lib.RnError("candidate slab overflow: more than cand_cap=%d anchors of one (page, class) passed the "
                           "score threshold; results would be inexact -- raise cand_cap" % cand_cap)


def filter_detections(boxes, classification, other=None, class_specific_filter=True, nms=True,
                      score_threshold=0.05, max_detections=300, nms_threshold=0.5,
                      pre_nms_top_k=0, cand_cap=None):
    """model/layers.py:177-264 for ONE image: boxes (N,4), classification (N,C) ->
    ``[boxes (M,4) f32, scores (M,) f32, labels (M,) i32, *other]`` padded with -1 (M = max_detections)."""
    other = [] if other is None else list(other)
    b = _cuda(boxes)[None]
    c = _cuda(classification)[None]
    ob, osc, ol, oi, status = _run_filter(b, c, class_specific_filter, nms, score_threshold, max_detections,
                                          nms_threshold, pre_nms_top_k, cand_cap)
    _raise_on_overflow(status, cand_cap)
    extra = _gather_other([_cuda(o, o.dtype if isinstance(o, torch.Tensor) else torch.float32)[None] for o in other], oi, max_detections)
    return [ob[0], osc[0], ol[0]] + [e[0] for e in extra]


class FilterDetections(_Layer):
    """model/layers.py:267-370.  ``FilterDetections(...)([boxes, classification, *other])`` ->
    ``[boxes (B,M,4), scores (B,M), labels (B,M) int32, *other]``.  ``parallel_iterations`` is kept for
    config compatibility; the whole batch is always one set of launches."""

    def __init__(self, nms=True, class_specific_filter=True, nms_threshold=0.5, score_threshold=0.05,
                 max_detections=300, parallel_iterations=32, pre_nms_top_k=0, cand_cap=None, **kwargs):
        self.nms = nms
        self.class_specific_filter = class_specific_filter
        self.nms_threshold = nms_threshold
        self.score_threshold = score_threshold
        self.max_detections = max_detections
        self.parallel_iterations = parallel_iterations
        self.pre_nms_top_k = pre_nms_top_k          # extension, off by default (not in the reference)
        self.cand_cap = cand_cap
        super().__init__(**kwargs)
        self.last_indices = None

    def forward(self, inputs, **kwargs):
        boxes, classification, other = inputs[0], inputs[1], list(inputs[2:])
        ob, osc, ol, oi, status = _run_filter(boxes, classification, self.class_specific_filter, self.nms,
                                              self.score_threshold, self.max_detections, self.nms_threshold,
                                              self.pre_nms_top_k, self.cand_cap)
        _raise_on_overflow(status, self.cand_cap)
        self.last_indices = oi
        return [ob, osc, ol] + _gather_other(other, oi, self.max_detections)

    def compute_output_shape(self, input_shape):
        return [
            (input_shape[0][0], self.max_detections, 4),
            (input_shape[1][0], self.max_detections),
            (input_shape[1][0], self.max_detections),
        ] + [tuple([input_shape[i][0], self.max_detections] + list(input_shape[i][2:])) for i in range(2, len(input_shape))]

    def compute_mask(self, inputs, mask=None):
        return (len(inputs) + 1) * [None]

    def get_config(self):
        config = super().get_config()
        config.update({
            'nms': self.nms,
            'class_specific_filter': self.class_specific_filter,
            'nms_threshold': self.nms_threshold,
            'score_threshold': self.score_threshold,
            'max_detections': self.max_detections,
            'parallel_iterations': self.parallel_iterations,
        })
        return config


class DetectionHead(_Layer):
    """The inference tail of ``retinanet_bbox`` (model/defineModel.py:329-350) as one fused call:
    Anchors (P3..P7) -> RegressBoxes -> ClipBoxes -> FilterDetections.

    ``head([image, regression, classification, *other])`` -> ``[boxes, scores, labels, *other]``.
    ``image`` may be the (B,H,W,C) tensor or just its shape; only the shape is used.  Like the reference's
    ``convert_model`` (model/utils.py:231), NMS is on unless ``applyNms=False``."""

    def __init__(self, applyNms=True, class_specific_filter=True, anchor_params=None, mean=None, std=None,
                 nms_threshold=0.5, score_threshold=0.05, max_detections=300, pyramid_levels=None,
                 pre_nms_top_k=0, cand_cap=None, name='retinanet-bbox', **kwargs):
        super().__init__(name=name)
        self.nms = applyNms
        self.class_specific_filter = class_specific_filter
        self.anchor_params = anchor_params
        self.pyramid_levels = pyramid_levels
        self.mean = _norm_arg('mean', mean, [0, 0, 0, 0])
        self.std = _norm_arg('std', std, [0.2, 0.2, 0.2, 0.2])
        self.nms_threshold = nms_threshold
        self.score_threshold = score_threshold
        self.max_detections = max_detections
        self.pre_nms_top_k = pre_nms_top_k
        self.cand_cap = cand_cap
        self._specs = {}
        self.last_indices = None

    def spec_for(self, image_hw):
        key = (int(image_hw[0]), int(image_hw[1]))
        spec = self._specs.get(key)
        if spec is None:
            spec = _anchors.make_spec(key + (3,), self.pyramid_levels, self.anchor_params, None)
            self._specs[key] = spec
        return spec

    def workspace_bytes(self, batch, image_hw, num_classes):
        """Size of the filter workspace for a (batch, image shape, classes) configuration (static-buffer callers)."""
        N = self.spec_for(image_hw).num_anchors
        cap = N if self.cand_cap is None else int(self.cand_cap)
        return int(_lib.load().rn_filter_workspace_bytes(int(batch), N, int(num_classes), int(bool(self.class_specific_filter)),
                                                         cap, int(self.max_detections)))

    def forward(self, inputs, check=True, out=None, workspace=None, **kwargs):
        image, regression, classification, other = inputs[0], inputs[1], inputs[2], list(inputs[3:])
        shape = _shape_of(image)
        hw = (int(shape[1]), int(shape[2])) if len(shape) == 4 else (int(shape[0]), int(shape[1]))
        decode = dict(spec=self.spec_for(hw), regression=regression, mean=self.mean, std=self.std, clip_hw=hw)
        ob, osc, ol, oi, status = _run_filter(None, classification, self.class_specific_filter, self.nms,
                                              self.score_threshold, self.max_detections, self.nms_threshold,
                                              self.pre_nms_top_k, self.cand_cap, decode=decode, out=out, workspace=workspace)
        if check:
            _raise_on_overflow(status, self.cand_cap)
        self.last_indices = oi
        return [ob, osc, ol] + _gather_other(other, oi, self.max_detections)


def non_max_suppression(boxes, scores, max_output_size, iou_threshold=0.5):
    """``tf.image.non_max_suppression`` (call site model/layers.py:211) for one box set on the GPU:
    returns the selected indices (int32 CUDA tensor, selection order)."""
    b, s = _cuda(boxes).reshape(-1, 4), _cuda(scores).reshape(-1)
    K, M = int(b.shape[0]), int(max_output_size)
    lib = _lib.load()
    out = torch.empty((M,), dtype=torch.int32, device=b.device)
    count = torch.zeros((1,), dtype=torch.int32, device=b.device)
    ws_bytes = int(lib.rn_nms_workspace_bytes(K, M))
    ws = _lib.scratch("nms", ws_bytes, b.device)
    _lib.check(lib.rn_nms(_lib.ptr(b), _lib.ptr(s), K, M, float(np.float32(iou_threshold)), _lib.ptr(out),
                          _lib.ptr(count), _lib.ptr(ws), ws_bytes, _lib.stream_ptr(b.device)), "rn_nms")
    return out[:int(count.item())]


custom_objects = {
    'RegressBoxes': RegressBoxes,
    'FilterDetections': FilterDetections,
    'Anchors': Anchors,
    'ClipBoxes': ClipBoxes,
}
