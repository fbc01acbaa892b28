"""Import-path shims (ait_b200.compat): the reference's own import statements resolve to the drop-ins.
CPU part: registration / signatures only.  GPU part: the usage of adaptive_image_transformer.py:5-51 and the
`model._C` functions with the pybind signatures of lib/model/csrc/vision.cpp:7-13."""
import inspect
import sys

import pytest
import torch


@pytest.fixture()
def shims():
    from ait_b200 import compat
    compat.uninstall()
    for n in list(sys.modules):
        if n == "model" or n.startswith("model.") or n == "transformer" or n.startswith("transformer."):
            del sys.modules[n]
    names = compat.install()
    yield names
    compat.uninstall()


def test_shims_register_the_reference_import_paths(shims):
    from model import _C                                    # lib/model/roi_layers/nms.py:3
    from model.roi_layers import ROIAlign, nms, roi_align   # faster_rcnn_coatt_transformer_sk.py:16-17
    from model.roi_layers.nms import nms as nms2            # `from .nms import nms`
    from model.system.Models import Transformer as T1       # faster_rcnn_coatt_transformer_sk.py:27
    from transformer.Models import Transformer as T2        # adaptive_image_transformer.py:3
    import ait_b200.roi_layers as rl
    from ait_b200.system.Models import Transformer
    assert T1 is Transformer and T2 is Transformer
    assert nms is rl.nms and nms2 is rl.nms and ROIAlign is rl.ROIAlign and roi_align is rl.roi_align
    # pybind signatures (csrc/nms.h:10, ROIAlign.h:12-17, :31-41): positional argument counts
    assert len(inspect.signature(_C.nms).parameters) == 3
    assert len(inspect.signature(_C.roi_align_forward).parameters) == 6
    assert len(inspect.signature(_C.roi_align_backward).parameters) == 10
    with pytest.raises(RuntimeError):
        _C.roi_pool_forward(None, None, 1.0, 7, 7)
    # empty input: the CPU long tensor of csrc/nms.h:17-18, no device needed
    out = _C.nms(torch.zeros(0, 4), torch.zeros(0), 0.7)
    assert out.dtype == torch.int64 and out.numel() == 0 and out.device.type == "cpu"


def test_shims_uninstall_restores_sys_modules():
    from ait_b200 import compat
    compat.uninstall()
    before = {n: sys.modules.get(n) for n in compat._NAMES}
    compat.install()
    assert "model._C" in sys.modules
    compat.uninstall()
    assert {n: sys.modules.get(n) for n in compat._NAMES} == before


@pytest.mark.gpu
def test_reference_demo_usage_runs_against_the_shims(shims):
    """What adaptive_image_transformer.py:5-51 does, through the reference's import path: same constructor arguments, same
    keyword call, same output shape.  The demo leaves the module in train mode (dropout 0.1 active, output random): the
    drop-in then runs its training step with dropout, like the reference; .eval() is the deterministic inference engine."""
    from transformer.Models import Transformer
    batch_size, num_props, channels = 4, 128, 1024
    props_feat = torch.rand(batch_size * num_props, channels, 7, 7).cuda()
    non_qry = torch.rand(batch_size, channels, 8, 8).cuda()
    AIT = Transformer(d_k=64, d_v=64, d_model=channels // 2, d_word_vec=channels // 2, d_inner=channels * 2,
                      n_position=8 * 8, n_layers=1, n_head=8, dropout=0.1)
    AIT = AIT.cuda()
    assert "Transformer" in repr(AIT) and "layer_stack" in repr(AIT)      # print(AIT) of the demo
    out_train = AIT(x_props=props_feat, x_query=non_qry)                   # the demo's own call: .train(), dropout active
    assert tuple(out_train.shape) == (batch_size * num_props, channels, 8, 8) and out_train.dtype == torch.float32
    assert torch.isfinite(out_train).all() and AIT.last_dropout_seed != 0
    out = AIT.eval()(x_props=props_feat, x_query=non_qry)
    assert tuple(out.shape) == (batch_size * num_props, channels, 8, 8) and out.dtype == torch.float32
    assert torch.isfinite(out).all()
    assert float((out_train.detach() - out).abs().max()) > 1e-2 * float(out.abs().max())


@pytest.mark.gpu
def test_model_C_functions_match_the_oracle(shims):
    from conftest import load_golden
    from model import _C
    from oracle import c_ops
    g = load_golden("roi_align_small.pt")
    gen = torch.Generator().manual_seed(g["seed"])
    feat = torch.randn(2, 8, 38, 63, generator=gen)
    out = _C.roi_align_forward(feat.cuda(), g["rois"].cuda(), 1.0 / 16.0, 7, 7, 0)
    torch.testing.assert_close(out.cpu(), g["out"], rtol=1e-5, atol=1e-6)
    grad = torch.randn(out.shape, generator=gen)
    gin = _C.roi_align_backward(grad.cuda(), g["rois"].cuda(), 1.0 / 16.0, 7, 7, 2, 8, 38, 63, 0)
    ref = c_ops.roi_align_backward(grad.numpy(), g["rois"].numpy(), 1.0 / 16.0, 7, 7, 2, 8, 38, 63, 0)
    torch.testing.assert_close(gin.cpu(), torch.from_numpy(ref).float(), rtol=1e-4, atol=1e-5)
    n = 500
    boxes = torch.rand(n, 2, generator=gen) * 400
    boxes = torch.cat([boxes, boxes + 20 + torch.rand(n, 2, generator=gen) * 200], 1)
    scores = torch.rand(n, generator=gen)
    keep = _C.nms(boxes.cuda(), scores.cuda(), 0.5).cpu()
    order = torch.argsort(scores, descending=True, stable=True)
    ref_keep = torch.from_numpy(c_ops.nms_sorted(boxes[order].numpy(), 0.5, False, 0).astype("int64"))
    assert torch.equal(keep, torch.sort(order[ref_keep])[0])
