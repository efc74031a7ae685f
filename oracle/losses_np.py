"""Oracle (TEST INFRASTRUCTURE): numpy restatement of ``model/losses.py`` of the reference.

PARITY UNPINNED: the reference evaluates these with TensorFlow ops and TF autodiff;
TensorFlow / Keras are not installed here and the reference has no golden vectors for them.
The formulas below follow ``model/losses.py:22-44`` (focal) and ``:67-90`` (smooth-L1) line by
line; the one third-party function on the path, ``K.binary_crossentropy``
(``model/losses.py:37``), is restated from the published Keras / tf.keras backend source:

  * ``bce="tf2"`` (default; tf.keras 2.0-2.x and standalone Keras 2.3, when the prediction
    tensor is not *directly* the output of a Sigmoid op -- here it passes through
    Reshape/Concatenate/GatherNd first)::

        p_c = clip(p, 1e-7, 1 - 1e-7)
        bce = -( t * log(p_c + 1e-7) + (1 - t) * log(1 - p_c + 1e-7) )

  * ``bce="logits"`` (standalone Keras <= 2.2)::

        p_c = clip(p, 1e-7, 1 - 1e-7);  z = log(p_c / (1 - p_c))
        bce = max(z, 0) - z * t + log1p(exp(-|z|))

The backward pass is what TF autodiff produces for those graphs: the gradient flows through
both the focal weight (un-clipped p) and the BCE term (zero outside the clip range, whose
ends pass gradient: ``clip_by_value`` masks with ``>=``/``<=``); ignored anchors get 0;
``abs`` has gradient ``sign`` (0 at 0); ``where`` routes the gradient of the selected branch.

All arithmetic runs in ``dtype`` (float32 to mirror the reference, float64 for a "truth"
variant used to bound rounding differences in the tests).
"""
import numpy as np

_EPS = 1e-7  # keras.backend.epsilon()


def _bce_and_grad(t, p, mode, dt):
    eps = dt(_EPS)
    one = dt(1)
    pc = np.clip(p, eps, one - eps)
    inside = ((p >= eps) & (p <= one - eps)).astype(dt)
    if mode == "tf2":
        a = pc + eps
        b = one - pc + eps
        bce = -(t * np.log(a) + (one - t) * np.log(b))
        dbce = -(t / a - (one - t) / b) * inside
    elif mode == "logits":
        z = np.log(pc / (one - pc))
        bce = np.maximum(z, dt(0)) - z * t + np.log1p(np.exp(-np.abs(z)))
        sig = one / (one + np.exp(-z))
        dbce = (sig - t) / (pc * (one - pc)) * inside
    else:
        raise ValueError("unknown bce mode %r" % (mode,))
    return bce.astype(dt), dbce.astype(dt)


def focal(alpha=0.25, gamma=2.0, bce="tf2", dtype=np.float32):
    """model/losses.py:5-46.  Returns ``f(y_true (B,N,C+1), y_pred (B,N,C)) -> loss`` or, with
    ``return_grad=True``, ``(loss, dloss/dy_pred)``.  ``normalizer`` overrides the batch-local
    ``max(1, #state==1)`` (used to model the multi-GPU global normaliser)."""
    dt = dtype

    def _focal(y_true, y_pred, return_grad=False, normalizer=None):
        y_true = np.asarray(y_true, dtype=dt)
        p = np.asarray(y_pred, dtype=dt)
        t = y_true[:, :, :-1]
        state = y_true[:, :, -1]
        keep = (state != -1)[:, :, None]                       # :27-29
        is_pos = (t == 1)
        a_t = np.where(is_pos, dt(alpha), dt(1) - dt(alpha))   # :32-33
        base = np.where(is_pos, dt(1) - p, p)                  # :34
        fw = a_t * np.power(base, dt(gamma))                   # :35
        ce, dce = _bce_and_grad(t, p, bce, dt)
        per_elem = fw * ce                                     # :37
        n_pos = dt(max(1.0, float(np.count_nonzero(state == 1)))) if normalizer is None else dt(normalizer)
        loss = dt(np.sum(np.where(keep, per_elem, dt(0)), dtype=np.float64)) / n_pos   # :40-44
        if not return_grad:
            return loss
        dbase = np.where(is_pos, dt(-1), dt(1))
        dfw = a_t * dt(gamma) * np.power(base, dt(gamma) - dt(1)) * dbase
        grad = np.where(keep, dfw * ce + fw * dce, dt(0)) / n_pos
        return loss, grad.astype(dt)

    return _focal


def smooth_l1(sigma=3.0, dtype=np.float32):
    """model/losses.py:49-91.  ``f(y_true (B,N,5), y_pred (B,N,4))``; only rows with state == 1
    contribute; normaliser is the number of such *rows* (``max(1, .)``)."""
    dt = dtype
    s2 = sigma ** 2

    def _smooth_l1(y_true, y_pred, return_grad=False, normalizer=None):
        y_true = np.asarray(y_true, dtype=dt)
        pred = np.asarray(y_pred, dtype=dt)
        target = y_true[:, :, :-1]
        state = y_true[:, :, -1]
        pos = (state == 1)[:, :, None]                          # :72-74
        d = pred - target
        ad = np.abs(d)
        quad = ad < dt(1.0 / s2)                                 # :80-85
        per_elem = np.where(quad, dt(0.5 * s2) * np.power(ad, dt(2)), ad - dt(0.5 / s2))
        n_pos = dt(max(1, int(np.count_nonzero(state == 1)))) if normalizer is None else dt(normalizer)
        loss = dt(np.sum(np.where(pos, per_elem, dt(0)), dtype=np.float64)) / n_pos   # :88-90
        if not return_grad:
            return loss
        g = np.where(quad, dt(s2) * ad, dt(1)) * np.sign(d)
        grad = np.where(pos, g, dt(0)) / n_pos
        return loss, grad.astype(dt)

    return _smooth_l1
