"""Import-path shims: make the reference's own import statements resolve to the libaitb200-backed drop-ins.

    import ait_b200.compat; ait_b200.compat.install()

    from model import _C                          # lib/model/roi_layers/nms.py:3, roi_align.py:8
    from model.roi_layers import ROIAlign, nms    # faster_rcnn_coatt_transformer_sk.py:16-17, rpn/proposal_layer.py:21
    from model.system.Models import Transformer   # faster_rcnn_coatt_transformer_sk.py:27
    from transformer.Models import Transformer    # adaptive_image_transformer.py:3

`model._C` here is a module object with the three hot-path functions of the reference's pybind11 extension
(lib/model/csrc/vision.cpp:7-13) and their signatures:

    nms(dets, scores, threshold) -> int64 Tensor                                           (csrc/nms.h:10-28)
    roi_align_forward(input, rois, spatial_scale, pooled_h, pooled_w, sampling_ratio)      (csrc/ROIAlign.h:12-29)
    roi_align_backward(grad, rois, spatial_scale, pooled_h, pooled_w, batch, channels, h, w, sampling_ratio)   (:31-46)

`roi_pool_forward` / `roi_pool_backward` raise: ROIPool is out of scope (every shipped cfg sets POOLING_MODE: align).

Two situations:
  * the reference tree is NOT importable (the usual deployment: only this package): synthetic `model`, `model.system`,
    `transformer` packages are registered in sys.modules;
  * the reference's `lib/` IS on sys.path (`_init_paths.py`): its real `model` package is kept -- rpn, faster_rcnn, utils keep
    working -- and only the hot-path modules are overlaid (`model._C`, `model.roi_layers[.nms|.roi_align]`,
    `model.system.Models`, `transformer.Models`), so `_fasterRCNN` picks up the drop-ins without a source edit.
`uninstall()` restores sys.modules.
"""
import importlib
import importlib.util
import sys
import types

_saved = None

_NAMES = ("model", "model._C", "model.roi_layers", "model.roi_layers.nms", "model.roi_layers.roi_align",
          "model.roi_layers.roi_pool", "model.system", "model.system.Models", "model.system.Layers",
          "model.system.SubLayers", "transformer", "transformer.Models", "transformer.Layers", "transformer.SubLayers")


def _make_C():
    import torch
    from . import ops, roi_layers

    m = types.ModuleType("model._C")
    m.__doc__ = "libaitb200-backed stand-in for the reference's pybind11 extension (lib/model/csrc/vision.cpp:7-13)"

    def nms(dets, scores, threshold):
        return roi_layers.nms(dets, scores, float(threshold))

    def roi_align_forward(input, rois, spatial_scale, pooled_height, pooled_width, sampling_ratio):
        if not input.is_cuda:
            raise RuntimeError("model._C.roi_align_forward (ait_b200): CUDA tensors only (no CPU path)")
        b, c, h, w = input.shape
        if rois.size(0) == 0:
            return input.new_empty((0, c, pooled_height, pooled_width))
        nhwc = ops.transpose_cs(input.reshape(b, c, h * w), to_channels_last=True).view(b, h, w, c)
        return ops.roi_align_forward(nhwc, rois, float(spatial_scale), int(pooled_height), int(pooled_width),
                                     int(sampling_ratio), token_major=False)

    def roi_align_backward(grad, rois, spatial_scale, pooled_height, pooled_width, batch_size, channels, height, width,
                           sampling_ratio):
        if not grad.is_cuda:
            raise RuntimeError("model._C.roi_align_backward (ait_b200): CUDA tensors only (the reference has no CPU "
                               "backward either, csrc/ROIAlign.h:44)")
        if rois.size(0) == 0:
            return torch.zeros((batch_size, channels, height, width), dtype=grad.dtype, device=grad.device)
        return ops.roi_align_backward(grad, rois, float(spatial_scale), int(pooled_height), int(pooled_width),
                                      int(batch_size), int(channels), int(height), int(width), int(sampling_ratio))

    def _no_pool(*a, **k):
        raise RuntimeError("ait_b200: ROIPool is not provided (out of scope: all shipped cfgs use POOLING_MODE: align)")

    m.nms, m.roi_align_forward, m.roi_align_backward = nms, roi_align_forward, roi_align_backward
    m.roi_pool_forward = m.roi_pool_backward = _no_pool
    return m


def _reference_model_package():
    """the reference's real `model` package if its lib/ is on sys.path (never our own synthetic one)."""
    mod = sys.modules.get("model")
    if mod is not None and not getattr(mod, "__ait_b200_shim__", False):
        return mod
    try:
        spec = importlib.util.find_spec("model")
    except (ImportError, ValueError):
        spec = None
    if spec is None or spec.submodule_search_locations is None:
        return None
    import os
    if not any(os.path.isdir(os.path.join(p, "roi_layers")) for p in spec.submodule_search_locations):
        return None                                   # some unrelated package called `model`
    return importlib.import_module("model")


def install():
    """Register the shims (idempotent).  Returns the list of overlaid module names."""
    global _saved
    if _saved is not None:
        return sorted(_saved)
    from . import roi_layers
    from .system import Layers, Models, SubLayers

    saved = {n: sys.modules.get(n) for n in _NAMES}
    ref = _reference_model_package()
    if ref is None:
        model = types.ModuleType("model")
        model.__path__ = []
        model.__ait_b200_shim__ = True
        system = types.ModuleType("model.system")
        system.__path__ = []
        sys.modules["model"] = model
        sys.modules["model.system"] = system
    else:
        model = ref
        try:
            system = importlib.import_module("model.system")
        except Exception:
            system = types.ModuleType("model.system")
            system.__path__ = []
            sys.modules["model.system"] = system
    C = _make_C()
    sys.modules["model._C"] = C
    model._C = C

    rl = types.ModuleType("model.roi_layers")
    rl.__path__ = []
    rl.nms, rl.roi_align, rl.ROIAlign, rl._ROIAlign = roi_layers.nms, roi_layers.roi_align, roi_layers.ROIAlign, roi_layers._ROIAlign

    class ROIPool:  # noqa: D401  (constructed only when POOLING_MODE == 'pool')
        def __init__(self, *a, **k):
            raise RuntimeError("ait_b200: ROIPool is not provided (out of scope: all shipped cfgs use POOLING_MODE: align)")

    rl.ROIPool, rl.roi_pool = ROIPool, C.roi_pool_forward
    rl.__all__ = ["nms", "roi_align", "ROIAlign", "roi_pool", "ROIPool"]
    for name in ("model.roi_layers", "model.roi_layers.nms", "model.roi_layers.roi_align", "model.roi_layers.roi_pool"):
        sys.modules[name] = rl
    model.roi_layers = rl

    for name, mod in (("Models", Models), ("Layers", Layers), ("SubLayers", SubLayers)):
        sys.modules["model.system." + name] = mod
        setattr(system, name, mod)
    model.system = system

    tr = types.ModuleType("transformer")
    tr.__path__ = []
    tr.Models, tr.Layers, tr.SubLayers = Models, Layers, SubLayers
    sys.modules["transformer"] = tr
    for name, mod in (("Models", Models), ("Layers", Layers), ("SubLayers", SubLayers)):
        sys.modules["transformer." + name] = mod
    _saved = saved
    return sorted(saved)


def uninstall():
    global _saved
    if _saved is None:
        return
    for n, mod in _saved.items():
        if mod is None:
            sys.modules.pop(n, None)
        else:
            sys.modules[n] = mod
    _saved = None
