"""Bridge between the reference's Keras/TensorFlow graph seams and this package's torch-hosted CUDA kernels.

UNTESTED HERE: TensorFlow and Keras are not installed in the build image and cannot be installed (no network), so nothing
in this module has ever been executed.  It is an import-guarded sketch of the one piece a maintainer of the reference has
to add to use the kernels from `RetinaNet.py` unchanged; everything else in the package is tested on the GPU.

Why a bridge is needed.  The reference hands its losses to Keras (``training_model.compile(loss={'regression':
losses.smooth_l1(), 'classification': losses.focal()}, ...)``, RetinaNet.py:125-131) and places its layers in a Keras graph
(``model/defineModel.py:15-24, 340-350``).  Keras calls those objects with TensorFlow tensors (symbolic in graph mode).
``retinanet_b200.losses.focal()`` / ``smooth_l1()`` and ``retinanet_b200.layers.*`` take **torch** CUDA tensors: they are
drop-ins for the functions' *signatures and arithmetic*, not objects Keras can trace.  The seam therefore needs

* a zero-copy hand-over of device memory: ``tf.experimental.dlpack.to_dlpack`` <-> ``torch.utils.dlpack.from_dlpack``
  (both frameworks allocate from the same CUDA context of the process);
* an eager escape from the TF graph: ``tf.py_function``;
* ``tf.custom_gradient`` returning the gradient K2 has already computed in the same pass (the kernel is forward + backward).

Stream ordering: TF and torch use different CUDA streams; the sketch synchronises the torch stream before handing the
result back (correct, not fast -- a production bridge would share one stream or exchange events).
"""

try:                                    # pragma: no cover - TensorFlow is absent in the build image
    import tensorflow as tf
except Exception:                       # noqa: BLE001
    tf = None

import torch
from torch.utils import dlpack as _torch_dlpack


def available():
    return tf is not None


def _require_tf():
    if tf is None:
        raise ImportError("tf_bridge needs TensorFlow (>= 2.2 for tf.experimental.dlpack); it is not installed here -- "
                          "this module is an untested sketch, see its docstring")


def tf_to_torch(x):
    """Eager TF GPU tensor -> torch CUDA tensor sharing the memory."""
    _require_tf()
    return _torch_dlpack.from_dlpack(tf.experimental.dlpack.to_dlpack(x))


def torch_to_tf(t):
    """torch CUDA tensor -> TF tensor sharing the memory (the torch stream is synchronised first)."""
    _require_tf()
    torch.cuda.current_stream(t.device).synchronize()
    return tf.experimental.dlpack.from_dlpack(_torch_dlpack.to_dlpack(t.contiguous()))


def keras_loss(which, **kwargs):
    """``compile(loss={'regression': keras_loss('smooth_l1'), 'classification': keras_loss('focal')})``: a Keras-callable
    ``f(y_true, y_pred) -> scalar`` whose forward runs ``rn_focal_fwd_bwd`` / ``rn_smooth_l1_fwd_bwd`` and whose gradient is the
    one the kernel stored in the same pass (reference: model/losses.py:5-46, :49-91)."""
    _require_tf()
    from . import losses as _losses
    functor = {'focal': _losses.focal, 'smooth_l1': _losses.smooth_l1}[which](**kwargs)

    @tf.custom_gradient
    def loss(y_true, y_pred):
        def run(yt, yp):
            p = tf_to_torch(yp).detach().requires_grad_(True)
            value = functor(tf_to_torch(yt), p)             # torch.autograd.Function: forward stashes the gradient
            value.backward()
            return torch_to_tf(value.detach().reshape(())), torch_to_tf(p.grad)
        value, grad = tf.py_function(run, [y_true, y_pred], [tf.float32, tf.float32])
        value.set_shape(())
        grad.set_shape(y_pred.shape)
        return value, (lambda upstream: (None, upstream * grad))
    loss.__name__ = '_' + which                             # the names custom_objects uses (model/defineModel.py:22-23)
    return loss


def keras_detection_head(**head_kwargs):
    """A ``keras.layers.Layer`` around ``layers.DetectionHead`` for ``retinanet_bbox`` (model/defineModel.py:329-350):
    ``layer([image, regression, classification]) -> [boxes, scores, labels]``."""
    _require_tf()
    from . import layers as _layers
    head = _layers.DetectionHead(**head_kwargs)
    M = int(head.max_detections)

    class FilterDetectionsB200(tf.keras.layers.Layer):
        def call(self, inputs, **kwargs):
            image, regression, classification = inputs[0], inputs[1], inputs[2]

            def run(img, reg, cls):
                out = head([tuple(img.shape), tf_to_torch(reg), tf_to_torch(cls)])
                return [torch_to_tf(t) for t in out]
            boxes, scores, labels = tf.py_function(run, [image, regression, classification], [tf.float32, tf.float32, tf.int32])
            b = regression.shape[0]
            boxes.set_shape((b, M, 4)); scores.set_shape((b, M)); labels.set_shape((b, M))
            return [boxes, scores, labels]

        def get_config(self):
            return dict(super().get_config(), **{k: getattr(head, k) for k in
                                                 ('nms_threshold', 'score_threshold', 'max_detections', 'class_specific_filter')})
    return FilterDetectionsB200(name='filtered_detections')
