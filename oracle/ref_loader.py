"""Load the *reference's own* numpy half, unmodified, from /root/reference.

TEST INFRASTRUCTURE ONLY.  Used by ``tests/golden/make_golden.py`` (to generate the
committed fixtures) and by the optional live cross-checks in ``tests/`` when
``/root/reference`` is present (it is absent on the GPU box; nothing that runs there may
call this).  Never imported by the product package.

The reference imports keras / tensorflow / keras_resnet / matplotlib at module scope;
none are installed here, so minimal stub modules are seeded into ``sys.modules`` first
(recipe: SURVEY.md §8c).  Only the numpy functions are ever executed:
``model/anchors.py`` (generate_anchors, shift, guess_shapes, anchors_for_shape,
compute_gt_annotations, anchor_targets_bbox, bbox_transform) and
``model/utils.py:180-211`` (compute_overlap).
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("RN_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "anchors.py"))


_STUB_NAMES = []


def _stub(name, **attrs):
    import importlib.machinery
    mod = types.ModuleType(name)
    mod.__spec__ = importlib.machinery.ModuleSpec(name, None)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    _STUB_NAMES.append(name)
    return mod


def _remove_stubs():
    """The reference modules keep their references; other libraries probing for tensorflow / keras
    (torch._dynamo does) must not find the stubs afterwards."""
    for name in _STUB_NAMES:
        sys.modules.pop(name, None)
    del _STUB_NAMES[:]


def _install_stubs():
    if "keras" in sys.modules and getattr(sys.modules["keras"], "_rn_stub", False):
        return

    class _Layer(object):
        def __init__(self, *a, **k):
            pass

        def get_config(self):
            return {}

    class _Anything(object):
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return self

        def __getattr__(self, item):
            return _Anything()

    backend = _stub("keras.backend", floatx=lambda: "float32",
                    image_data_format=lambda: "channels_last")
    layers = _stub("keras.layers", Layer=_Layer)
    initializers = _stub("keras.initializers", Initializer=object)
    models = _stub("keras.models", Model=_Anything, load_model=_Anything())
    callbacks = _stub("keras.callbacks", Callback=object)
    kutils = _stub("keras.utils")
    keras = _stub("keras", backend=backend, layers=layers, initializers=initializers,
                  models=models, callbacks=callbacks, utils=kutils, _rn_stub=True)
    keras.regularizers = _Anything()
    keras.optimizers = _Anything()

    tf_config = types.SimpleNamespace(list_physical_devices=lambda kind=None: [])
    _stub("tensorflow", config=tf_config, image=_Anything(), nn=_Anything())

    kr_models = _stub("keras_resnet.models")
    _stub("keras_resnet", models=kr_models, custom_objects={})
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            _stub("matplotlib")


def load_reference():
    """Returns (anchors_module, utils_module) of the reference, imported in the order that
    avoids its circular import (model.utils first)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib
    if "model.anchors" in sys.modules and "model.utils" in sys.modules:
        return sys.modules["model.anchors"], sys.modules["model.utils"]
    try:
        ref_utils = importlib.import_module("model.utils")
        ref_anchors = importlib.import_module("model.anchors")
    finally:
        _remove_stubs()
    return ref_anchors, ref_utils


def load_reference_generator():
    """The reference's ``csv_generator`` module (``Generator.filter_annotations`` / ``compute_inputs`` /
    ``compute_targets`` are plain numpy).  ``Shapes`` lives in the archived FasterRCNN tree; ``keras.utils.Sequence``
    is stubbed as ``object``.  Instances are made with ``Generator.__new__`` (the constructor reads a dataset)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    import importlib
    if "csv_generator" in sys.modules and hasattr(sys.modules["csv_generator"], "Generator"):
        return sys.modules["csv_generator"]
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    shapes_dir = os.path.join(REFERENCE_ROOT, "FasterRCNN")      # only for `Shapes`; searched last
    if shapes_dir not in sys.path:
        sys.path.append(shapes_dir)
    try:
        sys.modules["keras"].utils.Sequence = object
        importlib.import_module("model.utils")
        importlib.import_module("model.anchors")
        gen = importlib.import_module("csv_generator")
    finally:
        _remove_stubs()
    return gen
