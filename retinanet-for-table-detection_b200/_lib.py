"""ctypes binding of librn_b200.so (C-ABI declared in include/rn_b200.h).

The product path has NO CPU fallback: if the library is missing or a call fails, an exception is
raised.  PyTorch is used only as plumbing -- device allocations, streams, pinned staging buffers.
"""
import ctypes
import os
import threading
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_longlong, c_size_t, c_void_p

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RN_B200_LIB") or os.path.join(_HERE, "librn_b200.so")   # env: A/B builds for profiles/ sweeps

RN_BCE_TF2 = 0
RN_BCE_LOGITS = 1
RN_LOSS_SHARED_STATE = 1
RN_LOSS_NPOS_PEER_BOX = 2
RN_LOSS_FROM_LOGITS = 4
RN_LOSS_PEER_LAG1 = 8
RN_LOSS_PEER_PUBLISH = 16
RN_LOSS_PEER_LOSSES = 32
RN_MAX_WORLD = 16


class RnError(RuntimeError):
    """A librn_b200 call returned a negative status."""


_P = c_void_p            # device pointers travel as integers
_HI = POINTER(c_int)     # small host tables
_HF = POINTER(c_float)

# name -> (restype, argtypes); mirrors include/rn_b200.h line by line
SIGNATURES = {
    "rn_version": (c_int, []),
    "rn_last_error": (c_char_p, []),
    "rn_anchor_targets": (c_int, [_P, _HI, _HI, c_int, c_int, _P, c_longlong, _P, _P, _P, _P, c_int, c_int, c_int,
                                  c_float, c_float, _P, _P, _P, _P, _P, _P]),
    "rn_anchor_targets_ordered": (c_int, [_P, _HI, _HI, c_int, c_int, _P, c_longlong, _P, _P, _P, _P, c_int, c_int, c_int,
                                          c_float, c_float, _P, _P, _P, _P, _P, _P, _P]),
    "rn_anchor_targets_sparse": (c_int, [_P, _HI, _HI, c_int, c_int, c_longlong, _P, _P, _P, _P, c_int, c_int,
                                         c_float, c_float, _P, _P, _P, _P, _P, _P]),
    "rn_anchors_f64": (c_int, [_P, _HI, _HI, c_int, c_int, _P, _P]),
    "rn_compute_overlap": (c_int, [_P, c_longlong, _P, c_int, _P, _P]),
    "rn_bbox_transform": (c_int, [_P, _P, c_longlong, POINTER(c_double), POINTER(c_double), _P, _P]),
    "rn_loss_workspace_bytes": (c_size_t, []),
    "rn_count_positive": (c_int, [_P, c_longlong, c_int, _P, _P, c_size_t, _P]),
    "rn_focal_fwd_bwd": (c_int, [_P, _P, c_longlong, c_int, c_float, c_float, c_int, _P, _P, _P, _P, c_size_t, _P]),
    "rn_smooth_l1_fwd_bwd": (c_int, [_P, _P, c_longlong, c_float, _P, _P, _P, _P, c_size_t, _P]),
    "rn_loss_fwd_bwd": (c_int, [_P, _P, _P, _P, c_longlong, c_int, c_float, c_float, c_int, c_float,
                                _P, _P, _P, _P, c_int, _P, c_size_t, _P]),
    "rn_loss_fwd_bwd_levels": (c_int, [_P, _P, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_longlong), c_int, c_int, c_int,
                                       c_float, c_float, c_int, c_float, _P, _P, POINTER(c_void_p), POINTER(c_void_p),
                                       c_int, _P, c_size_t, _P]),
    "rn_anchors_f32": (c_int, [_P, _HI, _HI, c_int, c_int, c_int, _P, _P]),
    "rn_regress_boxes": (c_int, [_P, _P, c_longlong, _HF, _HF, _P, _P]),
    "rn_clip_boxes": (c_int, [_P, c_longlong, c_float, c_float, _P, _P]),
    "rn_filter_workspace_bytes": (c_size_t, [c_int, c_longlong, c_int, c_int, c_longlong, c_int]),
    "rn_filter_detections": (c_int, [_P, _P, c_int, c_longlong, c_int, c_int, c_int, c_float, c_float, c_int,
                                     c_int, c_longlong, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "rn_decode_filter_detections": (c_int, [_P, _HI, _HI, c_int, c_int, _P, _P, c_int, c_longlong, c_int,
                                            _HF, _HF, c_float, c_float, c_int, c_int, c_float, c_float, c_int,
                                            c_int, c_longlong, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "rn_gather_other": (c_int, [_P, _P, c_int, c_longlong, c_int, c_int, _P, _P]),
    "rn_nms_workspace_bytes": (c_size_t, [c_longlong, c_int]),
    "rn_nms": (c_int, [_P, _P, c_longlong, c_int, c_float, _P, _P, _P, c_size_t, _P]),
    "rn_debug_nms_timing": (c_int, [c_int]),
    "rn_debug_filter_stages": (c_int, [c_int]),
    "rn_rescale_cut": (c_int, [_P, _P, _P, c_int, c_int, c_float, _P, _P, _P]),
    "rn_preprocess_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "rn_preprocess_pages": (c_int, [_P, c_int, c_int, c_int, _P, _P, _P, c_size_t, _P]),
    "rn_peer_box_bytes": (c_size_t, []),
    "rn_peer_box_create": (c_int, [c_int, POINTER(c_void_p), c_void_p]),
    "rn_peer_box_open": (c_int, [c_void_p, POINTER(c_void_p)]),
    "rn_peer_box_close": (c_int, [c_void_p]),
    "rn_peer_box_destroy": (c_int, [c_void_p]),
    "rn_peer_box_connect": (c_int, [c_void_p, POINTER(c_void_p), c_int, c_int]),
    "rn_peer_box_bind": (c_int, [c_void_p, POINTER(c_void_p), c_int, c_int, _P, _P]),
    "rn_peer_box_set_timeout": (c_int, [c_void_p, c_double]),
    "rn_peer_box_status": (c_int, [c_void_p, POINTER(c_int), c_int]),
    "rn_peer_box_step": (c_int, [c_void_p, POINTER(ctypes.c_ulonglong)]),
    "rn_peer_publish": (c_int, [_P, c_void_p, POINTER(c_void_p), c_int, c_int, _P]),
}

_lib = None
_lock = threading.Lock()


def load():
    """Load (once) and return the ctypes library.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise RnError("librn_b200.so not found at %s -- build it with "
                          "`python retinanet-for-table-detection_b200/build.py` (there is no CPU fallback)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status, what):
    if status != 0:
        msg = load().rn_last_error()
        raise RnError("%s failed (%d): %s" % (what, status, msg.decode() if msg else "?"))


def require_cuda():
    if not torch.cuda.is_available():
        raise RnError("no CUDA device visible: the B200 path has no CPU fallback")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def host_ints(values):
    arr = np.ascontiguousarray(np.asarray(values, dtype=np.int32).ravel())
    return arr, arr.ctypes.data_as(_HI)


def host_floats(values):
    arr = np.ascontiguousarray(np.asarray(values, dtype=np.float32).ravel())
    return arr, arr.ctypes.data_as(_HF)


# ------------------------------------------------------------------------------------------------
# per-(device, stream, tag) scratch tensors, grown on demand and reused
# ------------------------------------------------------------------------------------------------
_scratch = {}


def scratch(tag, nbytes, device, zero_init=False):
    key = (tag, torch.device(device).index, torch.cuda.current_stream(device).cuda_stream)
    t = _scratch.get(key)
    if t is None or t.numel() < nbytes:
        alloc = max(int(nbytes), 256)
        t = (torch.zeros if zero_init else torch.empty)(alloc, dtype=torch.uint8, device=device)
        _scratch[key] = t
    return t


def loss_workspace(device):
    """Zero-filled once; the loss kernels leave their ticket counter zeroed."""
    n = int(load().rn_loss_workspace_bytes())
    return scratch("loss", n, device, zero_init=True), n


def reset_scratch():
    _scratch.clear()
