"""Drop-in for the hot-path functions of the reference's ``model/utils.py``:
``compute_overlap`` (:180-211), TF ``shift`` (:51-80) and ``bbox_transform_inv`` (:84-112).
Everything else in that file (image resize, drawing, model conversion) is outside the path."""
import numpy as np
import torch

from . import _lib
from .anchors import AnchorSpec


def _device():
    _lib.require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def _as_cuda(x, dtype, device):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x), dtype=_NP[dtype])).to(device)


_NP = {torch.float32: np.float32, torch.float64: np.float64}


def compute_overlap(boxes1, boxes2):
    """model/utils.py:180-211.  (M,4) x (G,4) -> (M,G) float32 IoU, no +1 pixel convention; the
    arithmetic runs in fp64 (callers cast to float64, model/anchors.py:109) and each value is rounded
    to float32 once.  numpy in -> numpy out, CUDA tensors in -> CUDA tensor out."""
    device = _device()
    as_numpy = not isinstance(boxes1, torch.Tensor)
    b1 = _as_cuda(boxes1, torch.float64, device).reshape(-1, 4)
    b2 = _as_cuda(boxes2, torch.float64, device).reshape(-1, 4)
    out = torch.zeros((b1.shape[0], b2.shape[0]), dtype=torch.float32, device=device)
    _lib.check(_lib.load().rn_compute_overlap(_lib.ptr(b1), b1.shape[0], _lib.ptr(b2), b2.shape[0],
                                              _lib.ptr(out), _lib.stream_ptr(device)), "rn_compute_overlap")
    return out.cpu().numpy() if as_numpy else out


def shift(shape, stride, anchors):
    """model/utils.py:51-80 (the TensorFlow version used by the Anchors layer): float32 centres
    ``(arange + 0.5) * stride`` added to float32 base anchors -> (H*W*A, 4) float32 CUDA tensor."""
    device = _device()
    base = np.asarray(anchors.detach().cpu().numpy() if isinstance(anchors, torch.Tensor) else anchors)
    spec = AnchorSpec([[int(shape[0]), int(shape[1])]], [int(stride)], base[None].astype(np.float64))
    base32 = torch.from_numpy(np.ascontiguousarray(base[None].astype(np.float32))).to(device)
    out = torch.empty((spec.num_anchors, 4), dtype=torch.float32, device=device)
    if spec.num_anchors:
        _, hw_p = _lib.host_ints(spec.level_hw)
        _, st_p = _lib.host_ints(spec.strides)
        _lib.check(_lib.load().rn_anchors_f32(_lib.ptr(base32), hw_p, st_p, 1, spec.per_cell, 1,
                                              _lib.ptr(out), _lib.stream_ptr(device)), "rn_anchors_f32")
    return out


def bbox_transform_inv(boxes, deltas, mean=None, std=None):
    """model/utils.py:84-112: ``x1 = b_x1 + (d0*std0 + mean0) * width`` ... in float32, evaluation
    order as written.  (B,N,4) in -> (B,N,4) out (numpy -> numpy, CUDA tensor -> CUDA tensor)."""
    if mean is None:
        mean = [0, 0, 0, 0]
    if std is None:
        std = [0.2, 0.2, 0.2, 0.2]
    device = _device()
    as_numpy = not isinstance(boxes, torch.Tensor)
    b = _as_cuda(boxes, torch.float32, device)
    d = _as_cuda(deltas, torch.float32, device)
    if b.shape != d.shape or b.shape[-1] != 4:
        raise ValueError("boxes and deltas must have the same (..., 4) shape")
    out = torch.empty_like(b)
    m_arr, mp = _lib.host_floats(np.asarray(mean, dtype=np.float64))
    s_arr, sp = _lib.host_floats(np.asarray(std, dtype=np.float64))
    _lib.check(_lib.load().rn_regress_boxes(_lib.ptr(b), _lib.ptr(d), b.numel() // 4, mp, sp,
                                            _lib.ptr(out), _lib.stream_ptr(device)), "rn_regress_boxes")
    return out.cpu().numpy() if as_numpy else out
