"""CPU: pin the oracle (oracle/) against the golden vectors generated from the reference
(tests/golden/make_golden.py) and, where /root/reference is present, against the reference live."""
import hashlib
import os

import numpy as np
import pytest
import torch

from conftest import golden_head, head_inputs, load_golden
from oracle import c_ops, head_oracle, ref_import

torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ------------------------------------------------------------------------------------------ NMS
def test_nms_oracle_matches_reference_golden_rpn_site():
    from ait_b200 import synth
    g = load_golden("nms_rpn_unit0.pt")
    boxes, scores = synth.rpn_outputs(0)
    order = np.argsort(-scores.double().numpy(), kind="stable")[:6000]
    b6 = boxes.numpy()[order]
    keep = c_ops.nms_sorted(b6, 0.7, ge=False)
    assert len(keep) == g["n_keep"]
    assert sha(keep.astype(np.int64)) == g["keep_sha"]
    assert np.array_equal(keep[:300], g["keep_first300"].numpy())
    # `>=` variant (the CPU reference's own rule) gives the same list here: no exact ties at 0.7
    assert np.array_equal(c_ops.nms_sorted(b6, 0.7, ge=True), keep)
    # full nms(dets, scores, thr) contract: ascending ORIGINAL indices
    keep_all = c_ops.nms(boxes.numpy(), scores.numpy(), 0.7)
    assert len(keep_all) == g["n_keep_all"] and sha(keep_all) == g["keep_all_sha"]
    rois, counts = head_oracle.propose_rois(boxes[None], scores[None], 6000, 300, 0.7)
    assert counts == [300] and torch.equal(rois[0], g["rois"])


def test_nms_tie_divergence_cuda_gt_vs_cpu_ge():
    """IoU == thr exactly: CUDA keeps (>), CPU suppresses (>=)  (nms.cu:60 vs nms_cpu.cpp:60)."""
    boxes = np.array([[0, 0, 9, 9], [0, 0, 9, 4], [20, 20, 29, 29]], dtype=np.float32)  # IoU(0,1) = 50/100
    assert list(c_ops.nms_sorted(boxes, 0.5, ge=False)) == [0, 1, 2]
    assert list(c_ops.nms_sorted(boxes, 0.5, ge=True)) == [0, 2]


def test_nms_edge_cases():
    assert len(c_ops.nms(np.zeros((0, 4), np.float32), np.zeros(0, np.float32), 0.7)) == 0
    one = np.array([[1, 2, 3, 4]], np.float32)
    assert list(c_ops.nms(one, np.array([0.3], np.float32), 0.7)) == [0]
    same = np.tile(np.array([[5, 5, 50, 50]], np.float32), (70, 1))        # spans two 64-blocks
    assert list(c_ops.nms(same, np.linspace(1, 0, 70).astype(np.float32), 0.7)) == [0]
    # equal scores: stable (lower index first)
    b = np.array([[0, 0, 10, 10], [0, 0, 10, 10], [100, 100, 110, 110]], np.float32)
    assert list(c_ops.nms(b, np.array([0.5, 0.5, 0.5], np.float32), 0.7)) == [0, 2]


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")
def test_nms_oracle_vs_reference_live():
    C = ref_import.load_ref_C()
    g = torch.Generator().manual_seed(5)
    for n in (1, 63, 64, 65, 500, 2000):
        xy = torch.rand(n, 2, generator=g) * 400
        wh = torch.rand(n, 2, generator=g) * 150 + 1
        boxes = torch.cat([xy, xy + wh], 1)
        scores = torch.rand(n, generator=g)
        for thr in (0.3, 0.7):
            ref = C.nms(boxes, scores, thr).numpy()
            assert np.array_equal(c_ops.nms(boxes.numpy(), scores.numpy(), thr, ge=True), ref)


# ------------------------------------------------------------------------------------------ ROIAlign
def test_roi_align_oracle_matches_reference_golden():
    g = load_golden("roi_align_small.pt")
    feat = torch.randn(2, 8, 38, 63, generator=torch.Generator().manual_seed(g["seed"]))
    out = c_ops.roi_align_forward(feat.numpy(), g["rois"].numpy(), 1 / 16.0, 7, 7, 0)
    np.testing.assert_allclose(out, g["out"].numpy(), rtol=1e-5, atol=1e-6)
    out2 = c_ops.roi_align_forward(feat.numpy(), g["rois"].numpy(), 1 / 16.0, 7, 7, 2)
    np.testing.assert_allclose(out2, g["out_sr2"].numpy(), rtol=1e-5, atol=1e-6)


def test_roi_align_oracle_vs_torchvision():
    """the legacy kernel == torchvision roi_align(sampling_ratio=0, aligned=False)."""
    tv = pytest.importorskip("torchvision")
    from ait_b200 import synth
    feat = synth.c4_map(3, channels=16)[None]
    rois = synth.random_rois(3, 32)
    ref = tv.ops.roi_align(feat, rois, (7, 7), 1 / 16.0, 0, False)
    out = c_ops.roi_align_forward(feat.numpy(), rois.numpy(), 1 / 16.0, 7, 7, 0)
    np.testing.assert_allclose(out, ref.numpy(), rtol=1e-5, atol=1e-6)


def test_roi_align_backward_oracle_is_adjoint_of_forward():
    """<forward(x), g> == <x, backward(g)> (linearity / adjointness), checked in float64."""
    g = torch.Generator().manual_seed(2)
    feat = torch.randn(1, 4, 38, 63, generator=g)
    rois = torch.tensor([[0, 17.0, 33.0, 411.0, 288.0], [0, 600.0, 10.0, 990.0, 590.0], [0, 5.0, 5.0, 9.0, 9.0]])
    grad = torch.randn(3, 4, 7, 7, generator=g)
    out = c_ops.roi_align_forward(feat.numpy(), rois.numpy(), 1 / 16.0, 7, 7, 0)
    gin = c_ops.roi_align_backward(grad.numpy(), rois.numpy(), 1 / 16.0, 7, 7, 1, 4, 38, 63, 0)
    lhs = float((out.astype(np.float64) * grad.numpy().astype(np.float64)).sum())
    rhs = float((feat.numpy().astype(np.float64) * gin).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(lhs))


# ------------------------------------------------------------------------------------------ AIT + head
def test_ait_oracle_matches_reference_golden():
    head, _ = golden_head()
    g = load_golden("ait_rand.pt")
    gen = torch.Generator().manual_seed(g["seed"])
    xp = torch.rand(6, 1024, 7, 7, generator=gen)
    xq = torch.rand(2, 1024, 8, 8, generator=gen)
    with torch.no_grad():
        out = head_oracle.ait_forward(head.transformer.state_dict(), xp, xq)
    torch.testing.assert_close(out[:, ::8], g["out_s"], rtol=1e-4, atol=1e-5)


def test_head_oracle_matches_reference_golden():
    head, g = golden_head()
    non_img, non_qry, rois = head_inputs(g["B"], g["P"])
    assert torch.equal(rois, g["rois"])
    with torch.no_grad():
        o = head_oracle.head_forward(head.state_dict(), non_img, non_qry, rois)
    torch.testing.assert_close(o["pooled"][:, ::16], g["pooled_s"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(o["ait_out"][:, ::16], g["ait_s"], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(o["sk_out"][:, ::16], g["sk_s"], rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(o["feat"], g["feat"], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(o["qfeat"], g["qfeat"], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(o["bbox_pred"], g["bbox_pred"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(o["cls_prob"], g["cls_prob"], rtol=0, atol=1e-4)
    assert float(g["cls_prob"].max() - g["cls_prob"].min()) > 0.5     # the gate is not vacuous


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")
def test_state_dict_loads_into_reference_strict():
    import sys
    sys.path.insert(0, ref_import.REF_ROOT)
    head, _ = golden_head()
    ref_import.ref_transformer().load_state_dict(head.transformer.state_dict(), strict=True)
    ref_import.ref_sknet().load_state_dict(head.sk.state_dict(), strict=True)
    ref_import.ref_layer4().load_state_dict(head.RCNN_top.state_dict(), strict=True)


def test_oracle_autograd_matches_reference_golden_gradients():
    """config 4: the oracle restatement differentiated by torch autograd reproduces the gradients of the
    unmodified reference Transformer (tests/golden/ait_grad.pt, made by tests/golden/make_golden_grad.py)."""
    import torch
    from conftest import load_golden
    from ait_b200 import synth
    from oracle import head_oracle
    gold = load_golden("ait_grad.pt")
    g = torch.Generator().manual_seed(13)
    xp = torch.rand(2, 1024, 7, 7, generator=g).requires_grad_()
    xq = torch.rand(1, 1024, 8, 8, generator=g).requires_grad_()
    gout = torch.randn(2, 1024, 8, 8, generator=g)
    head = synth.make_head(seed=0, calibrated=True, randomize_bn=True)
    sd = {k: v.detach().clone().requires_grad_("pos_table" not in k) for k, v in head.transformer.state_dict().items()}
    out = head_oracle.ait_forward(sd, xp, xq)
    out.backward(gout)
    assert torch.allclose(out.detach()[:, ::8], gold["out_s"], rtol=1e-4, atol=1e-5)
    assert torch.allclose(xp.grad[:, ::4], gold["grad_props_s"], rtol=1e-3, atol=1e-5)
    assert torch.allclose(xq.grad[:, ::4], gold["grad_query_s"], rtol=1e-3, atol=1e-5)
    assert len(gold["params"]) == 46
    for name, ref in gold["params"].items():
        gr = sd[name].grad.reshape(-1)
        sample = gr[::ref["stride"]][:ref["sample"].numel()]
        err = float((sample - ref["sample"]).norm() / ref["sample"].norm())
        assert err < 1e-3, (name, err)


def test_oracle_dropout_sites_match_reference_golden():
    """.train() WITH dropout: the oracle with the seeded masks of oracle/drop_masks.py injected at its ten sites reproduces
    output and gradients of the unmodified reference Transformer whose ten nn.Dropout instances were fed the same masks
    (tests/golden/ait_drop.pt, made by tests/golden/make_golden_drop.py)."""
    from ait_b200 import synth
    from oracle import drop_masks
    gold = load_golden("ait_drop.pt")
    bs, P = gold["bs"], gold["num_props"]
    g = torch.Generator().manual_seed(gold["seed"])
    xp = torch.rand(bs * P, 1024, 7, 7, generator=g).requires_grad_()
    xq = torch.rand(bs, 1024, 8, 8, generator=g).requires_grad_()
    gout = torch.randn(bs * P, 1024, 8, 8, generator=g)
    masks = drop_masks.make_masks(gold["mask_seed"], bs, P, gold["p"], gold["p_attn"])
    head = synth.make_head(seed=0, calibrated=True, randomize_bn=True)
    sd = {k: v.detach().clone().requires_grad_("pos_table" not in k) for k, v in head.transformer.state_dict().items()}
    out = head_oracle.ait_forward(sd, xp, xq, drop=masks)
    out.backward(gout)
    with torch.no_grad():
        plain = head_oracle.ait_forward(sd, xp, xq)
    assert float((plain[:, ::8] - gold["out_s"]).abs().max()) > 1e-2 * float(gold["out_s"].abs().max())   # the masks matter
    assert torch.allclose(out.detach()[:, ::8], gold["out_s"], rtol=1e-4, atol=1e-5)
    assert torch.allclose(xp.grad[:, ::4], gold["grad_props_s"], rtol=1e-3, atol=1e-5)
    assert torch.allclose(xq.grad[:, ::4], gold["grad_query_s"], rtol=1e-3, atol=1e-5)
    assert len(gold["params"]) == 46
    for name, ref in gold["params"].items():
        gr = sd[name].grad.reshape(-1)
        sample = gr[::ref["stride"]][:ref["sample"].numel()]
        err = float((sample - ref["sample"]).norm() / ref["sample"].norm())
        assert err < 1e-3, (name, err)


def _proposal_inputs(seed=21, B=2, A=9, H=19, W=31):
    import torch
    g = torch.Generator().manual_seed(seed)
    cls_prob = torch.rand(B, 2 * A, H, W, generator=g)
    bbox_pred = 0.3 * torch.randn(B, 4 * A, H, W, generator=g)
    im_info = torch.tensor([[300.0, 500.0, 1.5], [280.0, 480.0, 0.8]])[:B]
    return cls_prob, bbox_pred, im_info


def test_oracle_proposal_layer_matches_reference_golden():
    """row f1: the restated proposal layer (anchors, bbox_transform_inv, clip, sort, top-n, NMS, pad) reproduces the
    rois of the unmodified `_ProposalLayer` bit for bit (tests/golden/make_golden_proposal.py)."""
    import torch
    from conftest import load_golden
    from oracle import head_oracle
    gold = load_golden("proposal_layer.pt")
    cls_prob, bbox_pred, im_info = _proposal_inputs(gold["seed"])
    rois = head_oracle.proposal_layer(cls_prob, bbox_pred, im_info, gold["anchors"], 16, gold["pre"], gold["post"],
                                      gold["thr"])
    assert torch.equal(rois, gold["rois"])


def _rpn_inputs(seed=23, B=2, H=19, W=31):
    import torch
    g = torch.Generator().manual_seed(seed)
    base_feat = torch.randn(B, 1024, H, W, generator=g).relu()
    im_info = torch.tensor([[300.0, 500.0, 1.5], [280.0, 480.0, 0.8]])[:B]
    g = torch.Generator().manual_seed(seed + 1)
    sd = {"RPN_Conv.weight": torch.randn(512, 1024, 3, 3, generator=g) * 0.01,
          "RPN_Conv.bias": torch.randn(512, generator=g) * 0.1,
          "RPN_cls_score.weight": torch.randn(18, 512, 1, 1, generator=g) * 0.05,
          "RPN_cls_score.bias": torch.randn(18, generator=g) * 0.1,
          "RPN_bbox_pred.weight": torch.randn(36, 512, 1, 1, generator=g) * 0.01,
          "RPN_bbox_pred.bias": torch.randn(36, generator=g) * 0.05}
    return base_feat, im_info, sd


def test_oracle_rpn_head_matches_reference_golden():
    """row f3 (RPN head): conv 3x3 + ReLU, cls / bbox 1x1 heads, pair softmax, proposal layer -- the restatement
    reproduces the unmodified `_RPN` (tests/golden/make_golden_rpn.py): rois bit for bit."""
    import torch
    from conftest import load_golden
    from oracle import head_oracle
    gold = load_golden("rpn_head.pt")
    base_feat, im_info, sd = _rpn_inputs(gold["seed"])
    torch.set_num_threads(8)
    rois, prob, bbox = head_oracle.rpn_forward(sd, base_feat, im_info, gold["anchors"], 16, gold["pre"], gold["post"],
                                               gold["thr"])
    assert torch.allclose(prob, gold["cls_prob"], rtol=1e-5, atol=1e-6)
    assert torch.allclose(bbox[:, ::3], gold["bbox_pred_s"], rtol=1e-5, atol=1e-6)
    assert torch.equal(rois[..., 0], gold["rois"][..., 0])
    assert torch.allclose(rois, gold["rois"], rtol=1e-5, atol=1e-3)


def _coatt_inputs(seed=29, B=2, H=19, W=31):
    import torch
    g = torch.Generator().manual_seed(seed)
    x_img, x_qry = torch.randn(B, 1024, H, W, generator=g).relu(), torch.randn(B, 1024, 8, 8, generator=g).relu()
    g = torch.Generator().manual_seed(seed + 1)
    sd = {}
    for name in ("emb", "rho", "phi"):
        sd[name + ".weight"] = torch.randn(512, 1024, 1, 1, generator=g) * 0.03
        sd[name + ".bias"] = torch.randn(512, generator=g) * 0.1
    for name in ("omega", "theta"):
        sd[name + ".0.weight"] = torch.randn(1024, 512, 1, 1, generator=g) * 0.05
        sd[name + ".0.bias"] = torch.randn(1024, generator=g) * 0.1
        sd[name + ".1.weight"] = torch.rand(1024, generator=g) + 0.5
        sd[name + ".1.bias"] = torch.randn(1024, generator=g) * 0.2
    return x_img, x_qry, sd


def test_oracle_coattention_matches_reference_golden():
    """row f3 (co-attention): the restatement reproduces the unmodified `B.CoAttention` (make_golden_coatt.py)."""
    import torch
    from conftest import load_golden
    from oracle import head_oracle
    gold = load_golden("coattention.pt")
    x_img, x_qry, sd = _coatt_inputs(gold["seed"])
    assert sorted(sd) == gold["keys"]
    torch.set_num_threads(8)
    non_img, non_qry = head_oracle.coattention_forward(sd, x_img, x_qry)
    assert torch.allclose(non_img[:, ::32], gold["non_img_s"], rtol=1e-4, atol=1e-4)
    assert torch.allclose(non_qry[:, ::16], gold["non_qry_s"], rtol=1e-4, atol=1e-4)


def _target_inputs(gold):
    import torch
    from oracle import target_oracle as T
    B, A, H, W, R = gold["shape"]
    gt, nb = T.synth_gt_boxes(gold["seeds"]["gt"], B)
    rois = T.synth_rois(gold["seeds"]["rois"], B, R, gt)
    im_info = torch.tensor([[300.0, 500.0, 1.0]] * B)
    return gt, nb, rois, im_info


def _loss_inputs(gold):
    import torch
    B, A, H, W, R = gold["shape"]
    g = torch.Generator().manual_seed(gold["seeds"]["loss"])
    return (torch.randn(B, 2 * A, H, W, generator=g), 0.4 * torch.randn(B, 4 * A, H, W, generator=g),
            torch.randn(B * 128, 2, generator=g), 0.8 * torch.randn(B * 128, 4, generator=g))


def test_oracle_target_layers_match_reference_golden():
    """row f4: the restated `_AnchorTargetLayer` / `_ProposalTargetLayer` consume numpy's global stream like the
    unmodified reference and reproduce its outputs bit for bit (tests/golden/make_golden_targets.py)."""
    import numpy as np
    import torch
    from conftest import load_golden
    from oracle import target_oracle as T
    gold = load_golden("targets.pt")
    gt, nb, rois, im_info = _target_inputs(gold)
    B, A, H, W, R = gold["shape"]
    np.random.seed(gold["seeds"]["anchor_np"])
    out = T.anchor_target(gold["anchors"], H, W, 16, gt, im_info)
    for got, ref in zip(out, gold["anchor_target"]):
        assert torch.equal(got, ref)
    np.random.seed(gold["seeds"]["proposal_np"])
    out = T.proposal_target(rois, gt)
    for got, ref in zip(out, gold["proposal_target"]):
        assert torch.equal(got, ref)


def test_oracle_losses_match_reference_golden():
    """row f4: the five training losses and their gradients w.r.t. the network outputs."""
    import torch
    from conftest import load_golden
    from oracle import target_oracle as T
    gold = load_golden("targets.pt")
    B = gold["shape"][0]
    rpn_cls_score, rpn_bbox_pred, score, bbox_pred = [t.requires_grad_() for t in _loss_inputs(gold)]
    a, p = gold["anchor_target"], gold["proposal_target"]
    l_rc, l_rb = T.rpn_losses(rpn_cls_score, rpn_bbox_pred, *a)
    l_c, l_m, l_b = T.rcnn_losses(score, bbox_pred, p[1], p[2], p[3], p[4], B, margin=gold["margin"])
    for got, key in ((l_rc, "rpn_cls"), (l_rb, "rpn_box"), (l_c, "cls"), (l_m, "margin"), (l_b, "bbox")):
        assert torch.allclose(got, gold["losses"][key], rtol=1e-6, atol=0), key
    (l_rc + l_rb + l_c + l_m + l_b).backward()
    for t, key in ((rpn_cls_score, "rpn_cls_score"), (rpn_bbox_pred, "rpn_bbox_pred"), (score, "score"),
                   (bbox_pred, "bbox_pred")):
        assert torch.allclose(t.grad, gold["grads"][key], rtol=1e-5, atol=1e-9), key


def test_oracle_target_layers_vs_reference_live():
    """Other shapes / seeds against the reference itself where /root/reference exists (sub-sampling of both fg and
    bg anchors, images without foreground rois)."""
    import numpy as np
    import pytest
    import torch
    from oracle import ref_import, target_oracle as T
    if not ref_import.available():
        pytest.skip("reference tree not present")
    ref_import.install()
    from model.rpn.anchor_target_layer import _AnchorTargetLayer
    from model.rpn.proposal_target_layer_cascade import _ProposalTargetLayer
    from model.utils.config import cfg_from_file
    import os
    # the effective training configuration (trainval_net_voc.py:206-209): config.py merged with cfgs/res50.yml -> BG_THRESH_LO 0.0
    cfg_from_file(os.path.join(ref_import.REF_ROOT, "cfgs", "res50.yml"))
    at, pt = _AnchorTargetLayer(16, [8, 16, 32], [0.5, 1, 2]), _ProposalTargetLayer(2)
    for seed, B, H, W, imh, imw, n_max in ((1, 2, 38, 63, 600.0, 1000.0, 6), (2, 4, 25, 40, 400.0, 640.0, 14)):
        gt, nb = T.synth_gt_boxes(seed, B, im_h=imh, im_w=imw, n_min=1, n_max=n_max)
        if seed == 2:
            gt[1, :, :4] *= 0.1          # tiny boxes: few foreground rois / anchors in image 1
        rois = T.synth_rois(seed + 100, B, 300, gt, im_h=imh, im_w=imw)
        im_info = torch.tensor([[imh, imw, 1.0]] * B)
        np.random.seed(seed)
        ref = at((torch.zeros(B, 18, H, W), gt, im_info, nb))
        np.random.seed(seed)
        got = T.anchor_target(at._anchors, H, W, 16, gt, im_info)
        for g_, r_ in zip(got, ref):
            assert torch.equal(g_, r_)
        np.random.seed(seed + 1)
        ref = pt(rois, gt, nb)
        np.random.seed(seed + 1)
        got = T.proposal_target(rois, gt)
        for g_, r_ in zip(got, ref):
            assert torch.equal(g_, r_)


def test_oracle_whole_head_training_step_matches_reference_golden_gradients():
    """tests/golden/head_grad.pt: the three detection losses and the gradients of the pooled features, the query feature
    and all 70 trainable parameters produced by the UNMODIFIED reference modules (Transformer, SKNet, layer4 with frozen
    BatchNorm, the Linear heads, the loss lines of `_fasterRCNN.forward`; CPU fp32 autograd -- tests/golden/
    make_golden_head_grad.py).  The oracle's differentiable restatement (head_oracle + target_oracle, fp64) is the gradient
    oracle of the GPU training tests; here it is pinned: losses to 1e-5, every gradient to 2e-3 relative L2 on the
    committed samples and 2e-3 on the norm (fp32 reference vs fp64 oracle)."""
    import torch
    from conftest import load_golden
    from ait_b200 import synth
    from oracle import head_oracle, target_oracle
    gold = load_golden("head_grad.pt")
    B, P = gold["B"], gold["P"]
    head = synth.make_head(seed=0, calibrated=True, randomize_bn=True)
    pnames = {n for n, _ in head.named_parameters()}
    sd = {k: (v.clone().double().requires_grad_() if k in pnames else (v.clone().double() if v.is_floating_point() else v.clone()))
          for k, v in head.state_dict().items()}
    g = torch.Generator().manual_seed(gold["seed"])
    maps = torch.stack([synth.c4_map(u) for u in range(B)])
    qrys = torch.stack([synth.query_feat(u) for u in range(B)]).double().requires_grad_()
    rois = torch.stack([synth.random_rois(u, P, batch_index=u) for u in range(B)])
    label = torch.tensor([[1, 0, 0, 1], [0, 0, 1, 0]]).view(-1)
    tgt = 0.3 * torch.randn(B * P, 4, generator=g)
    inw = (label > 0).float().view(-1, 1).expand(-1, 4).contiguous()
    pooled = head_oracle.roi_align(maps, rois.view(-1, 5)).double().requires_grad_()      # the differentiated leaf, as in the golden
    ref = head_oracle.head_forward(sd, maps, qrys, rois, dtype=torch.float64, roi_align_fn=lambda f, r: pooled)
    losses = target_oracle.rcnn_losses(ref["score"], ref["bbox_pred"].view(-1, 4), label, tgt.double(), inw.double(),
                                       inw.double(), B)
    sum(losses).backward()
    for a, b in zip(losses, gold["losses"]):
        assert abs(float(a.detach()) - b) < 1e-5 * max(1.0, abs(b)), (float(a.detach()), b)
    assert torch.allclose(ref["score"].detach().float(), gold["score"], rtol=1e-4, atol=1e-5)

    def check(name, grad, ref_entry):
        assert grad is not None, name
        f = grad.reshape(-1)
        s = f[::ref_entry["stride"]][:ref_entry["sample"].numel()]
        r = ref_entry["sample"].double()
        assert float((s - r).norm() / r.norm()) < 2e-3, (name, float((s - r).norm() / r.norm()))
        assert abs(float(f.norm()) / ref_entry["norm"] - 1.0) < 2e-3, name

    check("pooled", pooled.grad, gold["grad_pooled"])
    check("query", qrys.grad, gold["grad_query"])
    assert len(gold["params"]) == 70
    for name, entry in gold["params"].items():
        check(name, sd[name].grad, entry)
    # the selective-kernel attention the reference's SKBlock.forward discards gets no gradient in the oracle either
    # (the other parameters missing from the golden are the frozen BatchNorm scales / shifts, requires_grad = False there)
    rest = pnames - set(gold["params"])
    assert all(".bn" in n or "downsample.1" in n or n.startswith("sk.") for n in rest), rest
    for name in rest:
        if name.startswith("sk."):
            assert sd[name].grad is None or float(sd[name].grad.abs().max()) == 0.0, name


def test_detection_postprocessing_oracle_matches_reference_golden():
    """Row f2 pin: tests/golden/detections.pt holds what the reference's OWN script lines (test_net_voc.py:380-450, exec'd
    unmodified by make_golden_detections.py with the reference's cfg / bbox_transform_inv / clip_boxes / nms) produce; the
    restatement `head_oracle.detections` must give the same detections in the same order."""
    g = load_golden("detections.pt")
    for case in g["cases"]:
        mine = head_oracle.detections(g["rois"], g["cls_prob"], g["bbox_pred"], g["im_info"], case["thresh"],
                                      g["nms_thresh"], case["max_per_image"])
        for b, ref in enumerate(case["dets"]):
            assert mine[b].shape == ref.shape, (case["thresh"], case["max_per_image"], b, mine[b].shape, ref.shape)
            assert torch.equal(mine[b][:, 4], ref[:, 4])
            torch.testing.assert_close(mine[b][:, :4], ref[:, :4], rtol=1e-6, atol=1e-5)
