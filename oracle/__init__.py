"""CPU oracle for the RetinaNet anchor + detection-head path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the *checker* (or the timed CPU
baseline) -- never as something the product path calls.  The product package
(``retinanet-for-table-detection_b200``) does not import it and raises when its CUDA
library is missing.

What it restates (file:line are relative to the reference tree, jabhinav/RetinaNet-for-Table-Detection):

* ``anchors_np``  -- numpy half: ``model/anchors.py`` + ``model/utils.py:180-211``.
  PARITY PINNED: checked bit-for-bit against the reference's own code executed in the
  build container (``oracle/ref_loader.py``), fixtures committed under ``tests/golden/``
  by ``tests/golden/make_golden.py``.
* ``losses_np``   -- ``model/losses.py:5-91`` (focal, smooth-L1) forward and the gradient
  TF autodiff would give.
* ``layers_np``   -- ``model/utils.py:51-112`` (TF shift, bbox_transform_inv) and
  ``model/layers.py`` (Anchors, RegressBoxes, ClipBoxes, filter_detections).
  PARITY UNPINNED for the TensorFlow half (``losses_np``, ``layers_np``): TensorFlow /
  Keras are un-vendored, un-pinned third-party dependencies of the reference, are not
  installed here, and the reference ships no golden vectors.  The restated third-party
  semantics (``tf.image.non_max_suppression``, ``tf.nn.top_k``, ``K.binary_crossentropy``)
  are written out in the docstrings of the functions that use them, and are cross-checked
  offline against ``torchvision.ops.nms`` / torch autograd in ``tests/``.
"""
