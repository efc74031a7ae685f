"""N4, the page-image producer of the reference (``DetectTablesUtils.py:183-262``: ``preProcessTrainValImages`` and
``preProcessSampleImages``) on the GPU: per page

    img = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    img = cv2.adaptiveThreshold(img, 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY, 11, 2)
    b, g, r = (cv2.distanceTransform(img, t, maskSize=5) for t in (cv2.DIST_L2, cv2.DIST_L1, cv2.DIST_C))
    cv2.imwrite(target, cv2.merge((b, g, r)))

becomes one call over a batch of decoded pages: ``preprocess_pages(bgr)`` returns the uint8 ``(B, H, W, 3)`` images that
``imwrite`` would encode (file decoding / encoding stay on the host, they are not arithmetic of the path).  Kernels:
``csrc/preprocess.cu`` behind ``rn_preprocess_pages``; bit-exact against OpenCV 4.13 for page widths that are a multiple of 8
(e.g. the reference's 2200 x 1712 pages), see ``oracle/preprocess_np.py``.  No CPU fallback.
"""
import numpy as np
import torch

from . import _lib


def preprocess_pages(bgr, return_binary=False, out=None):
    """``bgr``: uint8 ``(B, H, W, 3)`` or ``(H, W, 3)`` (numpy or torch, host or CUDA; BGR order as ``cv2.imread`` returns it).
    Returns the distance-transformed pages as a uint8 CUDA tensor of the same shape (channel 0: 5x5 chamfer "L2", 1: L1,
    2: chessboard; rounded, saturated at 255); with ``return_binary`` also the adaptive-threshold image ``(B, H, W)``."""
    _lib.require_cuda()
    lib = _lib.load()
    t = bgr if isinstance(bgr, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(bgr))
    if t.dtype != torch.uint8 or t.shape[-1] != 3 or t.dim() not in (3, 4):
        raise ValueError("expected uint8 pages of shape (B, H, W, 3) or (H, W, 3); got %s %s" % (t.dtype, tuple(t.shape)))
    single = t.dim() == 3
    if single:
        t = t[None]
    t = t.cuda().contiguous() if not t.is_cuda else t.contiguous()
    B, H, W = int(t.shape[0]), int(t.shape[1]), int(t.shape[2])
    dst = torch.empty_like(t) if out is None else out
    if tuple(dst.shape) != (B, H, W, 3) or dst.dtype != torch.uint8 or not dst.is_cuda:
        raise ValueError("out must be a uint8 CUDA tensor of shape %s" % ((B, H, W, 3),))
    binary = torch.empty((B, H, W), dtype=torch.uint8, device=t.device) if return_binary else None
    ws_bytes = int(lib.rn_preprocess_workspace_bytes(B, H, W))
    ws = _lib.scratch("preprocess", ws_bytes, t.device)
    _lib.check(lib.rn_preprocess_pages(_lib.ptr(t), B, H, W, _lib.ptr(dst), _lib.ptr(binary), _lib.ptr(ws), ws_bytes,
                                       _lib.stream_ptr(t.device)), "rn_preprocess_pages")
    if single:
        dst = dst[0]
        binary = binary[0] if binary is not None else None
    return (dst, binary) if return_binary else dst


def preprocess_page(img_bgr):
    """One page, numpy in / numpy out: what ``DetectTablesUtils.py:246-256`` writes for ``img_bgr`` (before encoding)."""
    return preprocess_pages(img_bgr).cpu().numpy()
