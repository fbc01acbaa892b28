"""CPU: the C-ABI library loads and exports every symbol include/aitb200.h declares; the mirror
modules keep the reference's state_dict keys; host-side helpers agree with the reference."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT, load_golden


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "aitb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(aitb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from ait_b200 import _lib
    from ait_b200.build import build
    build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    assert set(names) == set(_lib.SIGNATURES), "ctypes prototypes out of sync with the header"
    loaded = _lib.load(check_device=False)
    assert loaded.aitb_version() == 100
    assert loaded.aitb_head_workspace_bytes(8, 300, 0) > loaded.aitb_ait_workspace_bytes(8, 300, 0) > 0
    assert loaded.aitb_nms_workspace_bytes(8, 21546, 6000) > 8 * 6000 * 94 * 8


def test_no_product_import_of_oracle():
    """The product package must never route through the oracle (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "ait_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "oracle_ops" not in txt or f.endswith(".cu"), f


def test_cuda_path_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ait_b200 import _lib
    from ait_b200.roi_layers import ROIAlign, nms
    with pytest.raises(RuntimeError):
        _lib.load(check_device=True)
    with pytest.raises(RuntimeError):
        nms(torch.zeros(3, 4), torch.zeros(3), 0.5)
    with pytest.raises(RuntimeError):
        ROIAlign((7, 7), 1 / 16.0, 0)(torch.zeros(1, 8, 4, 4), torch.zeros(1, 5))
    # reference contract: empty dets -> empty CPU long tensor (csrc/nms.h:17-18)
    out = nms(torch.zeros(0, 4), torch.zeros(0), 0.5)
    assert out.dtype == torch.int64 and out.numel() == 0 and out.device.type == "cpu"


def test_state_dict_keys_match_reference():
    from ait_b200.head import DetectionHead
    keys = json.load(open(os.path.join(GOLDEN, "state_dict_keys.json")))
    head = DetectionHead()
    for name, mod in (("transformer", head.transformer), ("sk", head.sk), ("RCNN_top", head.RCNN_top)):
        mine = {k: list(v.shape) for k, v in mod.state_dict().items()}
        assert mine == keys[name], name
    assert len(keys["transformer"]) == 48 and sum(int(np.prod(s)) for k, s in keys["transformer"].items()
                                                   if "pos_table" not in k) == 8338944
    assert [k for k in head.state_dict() if k.startswith("RCNN_cls_score")] == [
        "RCNN_cls_score.0.weight", "RCNN_cls_score.0.bias", "RCNN_cls_score.1.weight", "RCNN_cls_score.1.bias"]


def test_transformer_envelope_is_enforced():
    from ait_b200.system.Models import Transformer
    with pytest.raises(NotImplementedError):
        Transformer(n_layers=6)                      # reference default; the engine supports n_layers=1
    with pytest.raises(AssertionError):
        Transformer(d_model=256, d_word_vec=512, n_layers=1)
    t = Transformer(d_k=64, d_v=64, d_model=512, d_word_vec=512, d_inner=2048, n_position=64, n_layers=1, n_head=8)
    with pytest.raises(RuntimeError):
        t.eval()(torch.zeros(4, 1024, 6, 6), torch.zeros(2, 1024, 8, 8))


def test_anchors_and_box_decoding_match_reference():
    from ait_b200 import proposal, synth
    g = load_golden("nms_rpn_unit0.pt")
    np.testing.assert_array_equal(proposal.generate_anchors(scales=(8, 16, 32)), np.array(g["anchors"]["voc"]))
    np.testing.assert_array_equal(proposal.generate_anchors(scales=(4, 8, 16, 32)), np.array(g["anchors"]["coco"]))
    a = proposal.shifted_anchors(38, 63)
    assert a.shape == (38 * 63 * 9, 4)
    assert torch.equal(a[9], torch.from_numpy(proposal.generate_anchors()).float()[0] + torch.tensor([16., 0, 16, 0]))
    boxes, scores = synth.rpn_outputs(0)
    assert boxes.shape == (21546, 4) and scores.unique().numel() == 21546
    assert boxes[:, 0::2].min() >= 0 and boxes[:, 0::2].max() <= 999 and boxes[:, 1::2].max() <= 599


def test_weight_packing_host_logic():
    from ait_b200.packing import HeadEngine, round_to_tf32
    x = torch.tensor([1.0, 1.0 + 2 ** -11, 1.0 + 2 ** -10, -3.14159265, 1e-30, 65504.0])
    r = round_to_tf32(x)
    assert torch.all((r.view(torch.int32) & 0x1FFF) == 0)
    assert torch.all((r - x).abs() <= x.abs() * 2 ** -11)
    assert r[1].item() == 1.0 + 2 ** -10            # ties away from zero, like cvt.rna.tf32.f32
    conv = torch.nn.Conv2d(4, 6, 3, padding=1, bias=False)
    bn = torch.nn.BatchNorm2d(6).eval()
    bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 2); bn.weight.data.normal_(); bn.bias.data.normal_()
    w, b = HeadEngine._fold_bn(conv, bn)
    xin = torch.randn(2, 4, 5, 5)
    torch.testing.assert_close(torch.nn.functional.conv2d(xin, w, b, padding=1), bn(conv(xin)), rtol=1e-5, atol=1e-5)
    tm = HeadEngine._tap_major(w)
    assert tm.shape == (6, 36) and torch.equal(tm[2, 4 * 5:4 * 6], w[2, :, 1, 2])   # tap (ky=1,kx=2)


def test_unit_sharding_is_a_partition():
    from ait_b200.sharding import shard_units
    for n_units in (1, 7, 8, 160):
        for world in (1, 2, 4, 8):
            parts = [shard_units(n_units, r, world) for r in range(world)]
            flat = [u for p in parts for u in p]
            assert flat == list(range(n_units))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_workspace_and_saved_buffer_accounting():
    """Host-only size queries of the C ABI (no device call): the training forward's saved-activation buffer holds the
    token-major pooled input at offset 0 and the token-major AIT result ([bp*64, 1024] fp32) entirely inside the buffer, at
    a 1024-byte-aligned offset that grows with the problem -- `Transformer(..., token_major_out=True)` hands out a view at
    that offset, so a wrong offset silently feeds the next stage the wrong tensor."""
    from ait_b200 import _lib
    lib = _lib.load(check_device=False)
    prev = 0
    for B, P in [(1, 1), (2, 3), (16, 128)]:
        total = int(lib.aitb_ait_saved_bytes(B, P))
        off_pooled = int(lib.aitb_ait_saved_offset(B, P, 0))
        off_ait = int(lib.aitb_ait_saved_offset(B, P, 1))
        assert off_pooled == 0
        assert off_ait % 1024 == 0 and off_ait >= B * P * 49 * 1024 * 4          # behind the pooled input at least
        assert off_ait + B * P * 64 * 1024 * 4 <= total
        assert off_ait > prev
        prev = off_ait
        assert int(lib.aitb_ait_backward_workspace_bytes(B, P)) > 0
    assert int(lib.aitb_ait_saved_offset(0, 4, 1)) == 0                            # bad sizes: 0, no crash
    for dt in (_lib.AITB_F32, _lib.AITB_BF16, _lib.AITB_F32S):
        assert int(lib.aitb_head_workspace_bytes(8, 300, dt)) > int(lib.aitb_head_workspace_bytes(1, 300, dt)) > 0


def test_training_weight_repacking_host_logic():
    """The dgrad weights of the training path are pure re-indexings of the packed forward weights (sk_train.dgrad_weights,
    top_train.dgrad_weight_3x3).  Checked on the CPU against torch autograd: a convolution of the output gradient with
    the re-packed weight (tap-major, the layout the GEMM kernels consume) must equal the input gradient of the forward
    convolution -- dense 3x3 (layer4 conv2) and grouped 1x1 / 3x3 (SKBlock, groups = 8)."""
    import torch
    import torch.nn.functional as F
    from ait_b200 import sk_train, top_train

    def tap_major(w):
        return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)

    def unpack(m, c_out, c_in, k):          # packed [c_out, k*k*c_in] -> conv weight [c_out, c_in, k, k]
        return m.view(c_out, k, k, c_in).permute(0, 3, 1, 2).contiguous()

    g = torch.Generator().manual_seed(3)
    # dense 3x3, 512 -> 512 on a 4x4 map
    w = torch.randn(512, 512, 3, 3, generator=g, dtype=torch.float64)
    x = torch.randn(2, 512, 4, 4, generator=g, dtype=torch.float64, requires_grad=True)
    gy = torch.randn(2, 512, 4, 4, generator=g, dtype=torch.float64)
    F.conv2d(x, w, padding=1).backward(gy)
    wd = top_train.dgrad_weight_3x3(tap_major(w), 512, 512)
    got = F.conv2d(gy, unpack(wd, 512, 512, 3), padding=1)
    assert torch.allclose(got, x.grad, rtol=1e-10, atol=1e-10)
    # grouped (8 x 128 -> 128) 1x1 and 3x3 on an 8x8 map
    w1 = torch.randn(1024, 128, 1, 1, generator=g, dtype=torch.float64)
    w3 = torch.randn(1024, 128, 3, 3, generator=g, dtype=torch.float64)
    x = torch.randn(1, 1024, 8, 8, generator=g, dtype=torch.float64, requires_grad=True)
    g1 = torch.randn(1, 1024, 8, 8, generator=g, dtype=torch.float64)
    g3 = torch.randn(1, 1024, 8, 8, generator=g, dtype=torch.float64)
    (F.conv2d(x, w1, groups=8) * g1).sum().add((F.conv2d(x, w3, padding=1, groups=8) * g3).sum()).backward()
    w1t, w3t = sk_train.dgrad_weights(tap_major(w1), tap_major(w3))
    got = F.conv2d(g1, unpack(w1t, 1024, 128, 1), groups=8) + F.conv2d(g3, unpack(w3t, 1024, 128, 3), padding=1, groups=8)
    assert torch.allclose(got, x.grad, rtol=1e-10, atol=1e-10)


def test_training_dropout_host_logic():
    """Host side of the training-mode dropout (no GPU): the reference's semantics -- p = `dropout` at the nn.Dropout sites,
    0.1 on the attention probabilities whatever the constructor says (Modules.py:9-14) --, `set_dropout`, the per-step seed
    drawn from torch's CPU generator (reproducible under torch.manual_seed, 0 when dropout is off), and the seeded mask
    helper of the oracle (Bernoulli(1 - p) multipliers, per-unit decoder-side sites)."""
    import torch
    from ait_b200.system.Models import Transformer, set_dropout
    from oracle import drop_masks
    t = Transformer(n_layers=1, dropout=0.3, n_position=64)
    assert {m.p_dropout for m in t.modules() if hasattr(m, "p_dropout")} == {0.3}
    assert {m.p_attn_dropout for m in t.modules() if hasattr(m, "p_attn_dropout")} == {0.1}
    torch.manual_seed(11)
    p, pa, s1 = t._draw_dropout()
    torch.manual_seed(11)
    assert t._draw_dropout() == (p, pa, s1) and (p, pa) == (0.3, 0.1) and s1 != 0 and t.last_dropout_seed == s1
    assert t._draw_dropout()[2] != s1
    set_dropout(t, 0.0, 0.0)
    assert t._draw_dropout() == (0.0, 0.0, 0)
    assert Transformer(n_layers=1, dropout=0.0, n_position=64, attn_dropout=0.0)._draw_dropout() == (0.0, 0.0, 0)
    t.encoder.layer_stack[0].pos_ffn.p_dropout = 0.5            # the engine has one probability per kind
    with pytest.raises(RuntimeError):
        t._draw_dropout()
    m = drop_masks.make_masks(5, 2, 3, 0.2, 0.1)
    assert set(m) == set(drop_masks.ROW_SITES) | set(drop_masks.ATTN_SITES)
    assert m["enc_ffn"].shape == (6 * 64, 512) and m["dec_emb"].shape == (2 * 64, 512)        # per pair / per unit
    assert m["enc_slf_attn"].shape == (6, 8, 64, 64) and m["dec_slf_attn"].shape == (2, 8, 64, 64)
    keep = m["enc_ffn"] > 0
    assert abs(float(keep.float().mean()) - 0.8) < 0.01 and torch.allclose(m["enc_ffn"][keep], torch.tensor(1.25))


def test_tiled_map_view_rows():
    """`ops.tiled_rows`: GEMM rows of the tiled H x W map view of the RPN 3x3 convolution (boxes of 128 positions)."""
    from ait_b200 import ops
    assert ops.tiled_rows(38, 63, 8) == 8 * 19 * 128        # 64 x 2 boxes, 19 per image
    assert ops.tiled_rows(19, 31, 2) == 2 * 10 * 128
    assert ops.tiled_rows(50, 100, 1) == 50 * 128           # wide maps: 128 x 1 boxes
