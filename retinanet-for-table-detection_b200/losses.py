"""Drop-in for the reference's ``model/losses.py``: ``focal(alpha, gamma)`` and ``smooth_l1(sigma)``
return functors ``f(y_true, y_pred) -> scalar`` (``RetinaNet.py:125-131``, ``defineModel.py:22-23``).

The functors are differentiable w.r.t. ``y_pred``: a ``torch.autograd.Function`` whose forward launches
kernel K2 (``csrc/losses.cu``), which produces the loss *and* the gradient in the same pass; backward
only scales the stored gradient by the incoming one.  :func:`detection_losses` runs both losses in a
single launch (the fused path the benchmark measures).

Extra keyword ``normalizer`` (not in the reference): a device float tensor holding the positive-anchor
COUNT to normalise with -- the per-page counts K1 returns, summed and, on several GPUs, all-reduced
(``distributed.global_positive_count``).  Without it the count is taken from ``y_true`` on the device
first, exactly as the reference does (``model/losses.py:40-42``, ``:88-89``).
"""
import ctypes

import numpy as np
import torch

from . import _lib

BCE_MODES = {"tf2": _lib.RN_BCE_TF2, "logits": _lib.RN_BCE_LOGITS}


def _prep(y_true, y_pred, last_true, last_pred):
    """Accept numpy or torch, return contiguous float32 CUDA tensors + whether inputs were numpy."""
    _lib.require_cuda()
    as_numpy = not isinstance(y_pred, torch.Tensor)
    if isinstance(y_pred, torch.Tensor) and y_pred.is_cuda:
        device = y_pred.device
    else:
        device = torch.device("cuda", torch.cuda.current_device())
    yt = torch.as_tensor(y_true).to(device=device, dtype=torch.float32).contiguous()
    yp = torch.as_tensor(y_pred).to(device=device, dtype=torch.float32)
    if not yp.is_contiguous():
        yp = yp.contiguous()
    if yt.shape[-1] != last_true(yp) or yt.shape[:-1] != yp.shape[:-1]:
        raise ValueError("y_true %s does not match y_pred %s" % (tuple(yt.shape), tuple(yp.shape)))
    if last_pred is not None and yp.shape[-1] != last_pred:
        raise ValueError("y_pred last dimension must be %d" % last_pred)
    return yt, yp, device, as_numpy


def _norm_tensor(normalizer, device):
    if normalizer is None:
        return None
    if isinstance(normalizer, torch.Tensor):
        n = normalizer.to(device=device, dtype=torch.float32).reshape(-1)
        return n.sum().reshape(1) if n.numel() != 1 else n.contiguous()
    return torch.tensor([float(normalizer)], dtype=torch.float32, device=device)


class _FocalFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_pred, y_true, npos, alpha, gamma, bce):
        device = y_pred.device
        R = y_pred.numel() // y_pred.shape[-1]
        C = y_pred.shape[-1]
        need_grad = ctx.needs_input_grad[0]
        loss = torch.empty((), dtype=torch.float32, device=device)
        grad = torch.empty_like(y_pred) if need_grad else None
        ws, ws_bytes = _lib.loss_workspace(device)
        _lib.check(_lib.load().rn_focal_fwd_bwd(_lib.ptr(y_true), _lib.ptr(y_pred), R, C, alpha, gamma, bce,
                                                _lib.ptr(npos), _lib.ptr(loss), _lib.ptr(grad),
                                                _lib.ptr(ws), ws_bytes, _lib.stream_ptr(device)), "rn_focal_fwd_bwd")
        ctx.grad = grad
        return loss

    @staticmethod
    def backward(ctx, g):
        return (ctx.grad * g if ctx.grad is not None else None), None, None, None, None, None


class _SmoothL1Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_pred, y_true, npos, sigma):
        device = y_pred.device
        R = y_pred.numel() // 4
        need_grad = ctx.needs_input_grad[0]
        loss = torch.empty((), dtype=torch.float32, device=device)
        grad = torch.empty_like(y_pred) if need_grad else None
        ws, ws_bytes = _lib.loss_workspace(device)
        _lib.check(_lib.load().rn_smooth_l1_fwd_bwd(_lib.ptr(y_true), _lib.ptr(y_pred), R, sigma,
                                                    _lib.ptr(npos), _lib.ptr(loss), _lib.ptr(grad),
                                                    _lib.ptr(ws), ws_bytes, _lib.stream_ptr(device)),
                   "rn_smooth_l1_fwd_bwd")
        ctx.grad = grad
        return loss

    @staticmethod
    def backward(ctx, g):
        return (ctx.grad * g if ctx.grad is not None else None), None, None, None


def focal(alpha=0.25, gamma=2.0, bce="tf2"):
    """model/losses.py:5-46.  ``bce`` selects the restated ``K.binary_crossentropy`` form ("tf2":
    clip + epsilon inside the logs; "logits": standalone Keras <= 2.2) -- see DESIGN.md."""
    if bce not in BCE_MODES:
        raise ValueError("bce must be one of %s" % sorted(BCE_MODES))

    def _focal(y_true, y_pred, normalizer=None):
        """y_true (B,N,C+1) [labels..., anchor state], y_pred (B,N,C) probabilities -> scalar."""
        yt, yp, device, as_numpy = _prep(y_true, y_pred, lambda p: p.shape[-1] + 1, None)
        out = _FocalFn.apply(yp, yt, _norm_tensor(normalizer, device), float(alpha), float(gamma), BCE_MODES[bce])
        return np.float32(out.item()) if as_numpy else out

    return _focal


def smooth_l1(sigma=3.0):
    """model/losses.py:49-91."""

    def _smooth_l1(y_true, y_pred, normalizer=None):
        """y_true (B,N,5) [4 targets, anchor state], y_pred (B,N,4) -> scalar."""
        yt, yp, device, as_numpy = _prep(y_true, y_pred, lambda p: 5, 4)
        out = _SmoothL1Fn.apply(yp, yt, _norm_tensor(normalizer, device), float(sigma))
        return np.float32(out.item()) if as_numpy else out

    return _smooth_l1


def detection_losses(y_true_reg, y_true_cls, reg_pred, cls_pred, normalizer=None,
                     alpha=0.25, gamma=2.0, sigma=3.0, bce="tf2", want_grads=True, out=None, workspace=None,
                     shared_state=False, peer_box=None, peer_lag=0, peer_publish=False, peer_losses=False):
    """Both losses, forward + backward, in ONE launch of K2 (``rn_loss_fwd_bwd``).

    All tensors are float32 CUDA: ``y_true_reg`` (B,N,5), ``y_true_cls`` (B,N,C+1) in the order
    ``anchor_targets_bbox`` returns them; ``reg_pred`` (B,N,4), ``cls_pred`` (B,N,C).
    Returns ``(losses, grad_cls, grad_reg)`` where ``losses`` is a 3-float device tensor
    ``[focal, smooth_l1, normaliser]`` and the gradients are d(loss)/d(pred) of the respective loss.

    ``shared_state=True``: the smooth-L1 part reads the anchor state from ``y_true_cls[..., -1]`` instead of
    ``y_true_reg[..., -1]``.  ``anchor_targets_bbox`` always writes the same state into both, so for its
    outputs the result is identical and 20 B/anchor of reads disappear; keep ``False`` for foreign tensors.

    ``peer_box``: a :class:`distributed.PeerCounter` whose ``publish`` was enqueued for this step on every rank;
    the kernel then takes the sum of the published counts as the normaliser (``normalizer`` is ignored);
    ``peer_lag=1`` selects the step published before the latest one (pipelined schedule, see ``pipeline``).
    ``peer_publish=True`` (``peer_box.bind(count)`` done once): fused publish -- this launch sends the rank's count
    itself and completes the step, ``publish`` is not called for it.
    ``peer_losses=True``: the two loss sums travel through the mailbox as well (the kernel's last CTA sends, collects and
    adds them in rank order), so ``losses`` is the loss of the whole merged batch on every rank -- what the reference's
    ``multi_gpu_model`` computes (``RetinaNet.py:106-112``, ``model/losses.py:44, :90``).  One such launch per step."""
    device = cls_pred.device
    C = cls_pred.shape[-1]
    R = cls_pred.numel() // C
    if out is None:
        losses = torch.empty(3, dtype=torch.float32, device=device)
        grad_cls = torch.empty_like(cls_pred) if want_grads else None
        grad_reg = torch.empty_like(reg_pred) if want_grads else None
    else:
        losses, grad_cls, grad_reg = out
    if workspace is None:
        ws, ws_bytes = _lib.loss_workspace(device)
    else:
        ws, ws_bytes = workspace, workspace.numel()       # caller-owned, zero-filled uint8 tensor
    flags = _lib.RN_LOSS_SHARED_STATE if shared_state else 0
    if peer_box is not None:                              # normaliser = sum of the counts the ranks published
        npos_ptr = ctypes.c_void_p(peer_box.box)
        flags |= _lib.RN_LOSS_NPOS_PEER_BOX | (_lib.RN_LOSS_PEER_LAG1 if peer_lag else 0)
        if peer_publish:
            if peer_lag or peer_box.bound is None:
                raise ValueError("peer_publish needs peer_box.bind(count) and is not combinable with peer_lag")
            flags |= _lib.RN_LOSS_PEER_PUBLISH
        if peer_losses:
            flags |= _lib.RN_LOSS_PEER_LOSSES
    else:
        if peer_losses:
            raise ValueError("peer_losses needs peer_box")
        npos = _norm_tensor(normalizer, device)
        npos_ptr = _lib.ptr(npos)
    _lib.check(_lib.load().rn_loss_fwd_bwd(_lib.ptr(y_true_cls), _lib.ptr(cls_pred), _lib.ptr(y_true_reg),
                                           _lib.ptr(reg_pred), R, C, float(alpha), float(gamma), BCE_MODES[bce],
                                           float(sigma), npos_ptr,
                                           _lib.ptr(losses), _lib.ptr(grad_cls), _lib.ptr(grad_reg),
                                           flags,
                                           _lib.ptr(ws), ws_bytes, _lib.stream_ptr(device)), "rn_loss_fwd_bwd")
    return losses, grad_cls, grad_reg


def detection_losses_levels(y_true_reg, y_true_cls, reg_levels, cls_levels, normalizer=None, from_logits=True,
                            alpha=0.25, gamma=2.0, sigma=3.0, bce="tf2", out=None, workspace=None, peer_box=None):
    """The fused losses fed by the heads' per-level outputs: ``cls_levels[l]`` is the
    (B, n_l, 1) classification tensor of pyramid level l -- LOGITS when ``from_logits`` (the heads'
    ``Activation('sigmoid')``, model/defineModel.py:123, is fused into the kernel), else probabilities -- and
    ``reg_levels[l]`` the (B, n_l, 4) regression; the ``Concatenate(axis=1)`` of model/defineModel.py:217 never
    happens.  Targets are the concatenated ``(B, N, .)`` tensors of ``anchor_targets_bbox`` (state shared).

    Returns ``(losses[3], grad_cls_levels, grad_reg_levels)``; the classification gradients are w.r.t. what was
    passed in (logits or probabilities).  C == 1, gamma == 2, TF2 cross-entropy only (``rn_loss_fwd_bwd_levels``)."""
    device = y_true_cls.device
    L = len(cls_levels)
    if len(reg_levels) != L or L < 1:
        raise ValueError("cls_levels and reg_levels must have the same, non-zero number of levels")
    B = int(y_true_cls.shape[0])
    rows = [int(c.shape[1]) for c in cls_levels]
    for c, r in zip(cls_levels, reg_levels):
        if tuple(c.shape) != (B, c.shape[1], 1) or tuple(r.shape) != (B, c.shape[1], 4) or not (c.is_contiguous() and r.is_contiguous()):
            raise ValueError("level tensors must be contiguous (B, n_l, 1) / (B, n_l, 4); got %s / %s" % (tuple(c.shape), tuple(r.shape)))
    if sum(rows) != int(y_true_cls.shape[1]) or tuple(y_true_reg.shape) != (B, sum(rows), 5) or int(y_true_cls.shape[2]) != 2:
        raise ValueError("targets %s / %s do not match the levels (%d anchors, 1 class)"
                         % (tuple(y_true_reg.shape), tuple(y_true_cls.shape), sum(rows)))
    if out is None:
        losses = torch.empty(3, dtype=torch.float32, device=device)
        g_cls = [torch.empty_like(c) for c in cls_levels]
        g_reg = [torch.empty_like(r) for r in reg_levels]
    else:
        losses, g_cls, g_reg = out
    if workspace is None:
        ws, ws_bytes = _lib.loss_workspace(device)
    else:
        ws, ws_bytes = workspace, workspace.numel()
    flags = _lib.RN_LOSS_SHARED_STATE | (_lib.RN_LOSS_FROM_LOGITS if from_logits else 0)
    if peer_box is not None:
        npos_ptr = ctypes.c_void_p(peer_box.box)
        flags |= _lib.RN_LOSS_NPOS_PEER_BOX
    else:
        npos = _norm_tensor(normalizer, device)
        npos_ptr = _lib.ptr(npos)
    arr = lambda ts: (ctypes.c_void_p * L)(*[t.data_ptr() for t in ts])
    rows_arr = (ctypes.c_longlong * L)(*rows)
    _lib.check(_lib.load().rn_loss_fwd_bwd_levels(_lib.ptr(y_true_cls), _lib.ptr(y_true_reg), arr(cls_levels), arr(reg_levels),
                                                  rows_arr, L, B, 1, float(alpha), float(gamma), BCE_MODES[bce], float(sigma),
                                                  npos_ptr, _lib.ptr(losses), arr(g_cls), arr(g_reg), flags,
                                                  _lib.ptr(ws), ws_bytes, _lib.stream_ptr(device)), "rn_loss_fwd_bwd_levels")
    return losses, g_cls, g_reg


class _DetectionLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cls_pred, reg_pred, y_true_cls, y_true_reg, npos, alpha, gamma, sigma, bce):
        want = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        losses, gc, gr = detection_losses(y_true_reg, y_true_cls, reg_pred, cls_pred, npos, alpha, gamma, sigma,
                                          bce, want_grads=want)
        ctx.gc, ctx.gr = gc, gr
        return losses[0], losses[1]

    @staticmethod
    def backward(ctx, g_focal, g_sl1):
        gc = ctx.gc * g_focal if ctx.gc is not None else None
        gr = ctx.gr * g_sl1 if ctx.gr is not None else None
        return gc, gr, None, None, None, None, None, None, None


def detection_loss(y_true_reg, y_true_cls, reg_pred, cls_pred, normalizer=None,
                   alpha=0.25, gamma=2.0, sigma=3.0, bce="tf2"):
    """Autograd-aware fused version: returns ``(focal_loss, smooth_l1_loss)`` scalars."""
    return _DetectionLossFn.apply(cls_pred.contiguous(), reg_pred.contiguous(), y_true_cls.contiguous(),
                                  y_true_reg.contiguous(), _norm_tensor(normalizer, cls_pred.device),
                                  float(alpha), float(gamma), float(sigma), bce)
