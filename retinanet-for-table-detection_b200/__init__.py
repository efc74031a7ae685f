"""retinanet-for-table-detection_b200 -- B200 (sm_100a) implementation of RetinaNet's anchor and
detection-head path, drop-in for the functions and layers of the reference's ``model/`` package
(jabhinav/RetinaNet-for-Table-Detection):

    from retinanet_b200 import anchors, utils, losses, layers       # <- from model import ...

``anchors``  anchors_for_shape, anchor_targets_bbox, compute_gt_annotations, bbox_transform, ...  (K1)
``utils``    compute_overlap, bbox_transform_inv, shift
``losses``   focal, smooth_l1 (+ fused detection_losses)                                           (K2)
``layers``   Anchors, RegressBoxes, ClipBoxes, FilterDetections, filter_detections, DetectionHead  (K3-K5)
``generator``   filter_annotations, compute_inputs, compute_targets   (csv_generator.py batching around K1)
``postprocess`` rescale_and_cut, read_annotations_csv, write_detections_csv  (RetinaNet.py post-step, CSV formats)
``preprocess``  preprocess_pages (DetectTablesUtils.py: grey -> adaptive threshold -> three distance transforms -> uint8 page)

All arithmetic runs in hand-written CUDA kernels behind the C-ABI of ``include/rn_b200.h``
(``librn_b200.so``, built by ``build.py``); there is no CPU fallback.
"""
from . import _lib  # noqa: F401
from . import anchors, distributed, generator, layers, losses, pipeline, postprocess, preprocess, utils  # noqa: F401
from .anchors import (AnchorParameters, AnchorParameters_default, anchor_targets_bbox,  # noqa: F401
                      anchors_for_shape, bbox_transform, compute_gt_annotations, generate_anchors, guess_shapes)
from .layers import (Anchors, ClipBoxes, DetectionHead, FilterDetections, RegressBoxes,  # noqa: F401
                     custom_objects, filter_detections)
from .losses import detection_loss, detection_losses, detection_losses_levels, focal, smooth_l1  # noqa: F401
from .utils import bbox_transform_inv, compute_overlap  # noqa: F401

__version__ = "0.1.0"
