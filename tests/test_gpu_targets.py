"""Row f4 on the device: anchor / proposal target layers and the training losses against the oracle
(oracle/target_oracle.py, pinned to the unmodified reference) and the reference's golden outputs."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _inputs(gold):
    from oracle import target_oracle as T
    B, A, H, W, R = gold["shape"]
    gt, nb = T.synth_gt_boxes(gold["seeds"]["gt"], B)
    rois = T.synth_rois(gold["seeds"]["rois"], B, R, gt)
    im_info = torch.tensor([[300.0, 500.0, 1.0]] * B)
    return gt, nb, rois, im_info


def _loss_inputs(gold):
    B, A, H, W, R = gold["shape"]
    g = torch.Generator().manual_seed(gold["seeds"]["loss"])
    return (torch.randn(B, 2 * A, H, W, generator=g), 0.4 * torch.randn(B, 4 * A, H, W, generator=g),
            torch.randn(B * 128, 2, generator=g), 0.8 * torch.randn(B * 128, 4, generator=g))


def _check_anchor(got, ref):
    assert torch.equal(got[0].cpu(), ref[0]), "anchor labels differ"
    assert torch.allclose(got[1].cpu(), ref[1], rtol=1e-5, atol=1e-6)      # logf vs host libm
    assert torch.equal(got[2].cpu(), ref[2])
    assert torch.equal(got[3].cpu(), ref[3])


def _check_proposal(got, ref):
    assert torch.equal(got[0].cpu(), ref[0]), "sampled rois differ"
    assert torch.equal(got[1].cpu(), ref[1]), "roi labels differ"
    assert torch.allclose(got[2].cpu(), ref[2], rtol=1e-5, atol=1e-5)
    assert torch.equal(got[3].cpu(), ref[3])
    assert torch.equal(got[4].cpu(), ref[4])


def test_target_layers_match_reference_golden():
    """Same numpy seed -> the same sampled anchors / rois as the unmodified reference (labels, rois, weights bit
    exact; regression targets to an ulp of log)."""
    from ait_b200.targets import AnchorTargetLayer, ProposalTargetLayer
    gold = load_golden("targets.pt")
    gt, nb, rois, im_info = _inputs(gold)
    B, A, H, W, R = gold["shape"]
    at = AnchorTargetLayer(16, [8, 16, 32], [0.5, 1, 2]).to(DEV)
    assert torch.equal(at._anchors.cpu(), gold["anchors"])
    np.random.seed(gold["seeds"]["anchor_np"])
    out = at((torch.zeros(B, 2 * A, H, W, device=DEV), gt.to(DEV), im_info.to(DEV), nb.to(DEV)))
    _check_anchor(out, gold["anchor_target"])
    pt = ProposalTargetLayer(2)
    np.random.seed(gold["seeds"]["proposal_np"])
    out = pt(rois.to(DEV), gt.to(DEV), nb.to(DEV))
    _check_proposal(out, gold["proposal_target"])
    assert int(pt.last_bad_flag.item()) == 0


@pytest.mark.parametrize("seed,B,H,W,imh,imw,n_max,R", [(1, 2, 38, 63, 600.0, 1000.0, 6, 2000), (2, 4, 25, 40, 400.0, 640.0, 14, 300),
                                                       (3, 1, 38, 63, 600.0, 1000.0, 20, 64), (4, 16, 38, 63, 600.0, 1000.0, 8, 2000)])
def test_target_layers_match_oracle(seed, B, H, W, imh, imw, n_max, R):
    """VOC-size maps (21 546 anchors), COCO-style crowded images, images with almost no foreground, fg and bg
    sub-sampling both active, the proposal layer's 2000 training rois."""
    from ait_b200.targets import AnchorTargetLayer, ProposalTargetLayer
    from ait_b200.proposal import generate_anchors
    from oracle import target_oracle as T
    gt, nb = T.synth_gt_boxes(seed, B, im_h=imh, im_w=imw, n_min=1, n_max=n_max)
    if seed == 2:
        gt[1, :, :4] *= 0.1
    rois = T.synth_rois(seed + 100, B, R, gt, im_h=imh, im_w=imw)
    im_info = torch.tensor([[imh, imw, 1.0]] * B)
    base = torch.from_numpy(generate_anchors(scales=np.array([8, 16, 32]), ratios=np.array([0.5, 1, 2]))).float()
    stage = {}
    np.random.seed(seed)
    ref = T.anchor_target(base, H, W, 16, gt, im_info, stage=stage)
    at = AnchorTargetLayer(16, [8, 16, 32], [0.5, 1, 2]).to(DEV)
    np.random.seed(seed)
    out = at((torch.zeros(B, 18, H, W, device=DEV), gt.to(DEV), im_info.to(DEV), nb.to(DEV)))
    _check_anchor(out, ref)
    n_fg = (ref[0] == 1).flatten(1).sum(1)
    assert int(n_fg.max()) <= 128 and int(((ref[0] >= 0).flatten(1).sum(1)).max()) <= 256
    np.random.seed(seed + 1)
    ref = T.proposal_target(rois, gt)
    pt = ProposalTargetLayer(2)
    np.random.seed(seed + 1)
    out = pt(rois.to(DEV), gt.to(DEV), nb.to(DEV))
    _check_proposal(out, ref)


def test_private_generator_sampling_properties():
    """rng=np.random.Generator: an independent stream, same invariants: <= 128 fg and <= 256 labelled anchors per image,
    sampled anchors are a subset of the candidates, exactly 128 rois with <= 32 foreground first."""
    from ait_b200.targets import AnchorTargetLayer, ProposalTargetLayer
    from ait_b200.proposal import generate_anchors
    from oracle import target_oracle as T
    B, H, W = 4, 38, 63
    gt, nb = T.synth_gt_boxes(9, B, im_h=600.0, im_w=1000.0, n_min=3, n_max=12)
    rois = T.synth_rois(10, B, 2000, gt, im_h=600.0, im_w=1000.0)
    im_info = torch.tensor([[600.0, 1000.0, 1.0]] * B)
    base = torch.from_numpy(generate_anchors(scales=np.array([8, 16, 32]), ratios=np.array([0.5, 1, 2]))).float()
    stage = {}
    T.anchor_target(base, H, W, 16, gt, im_info, stage=stage)
    pre = torch.full((B, H * W * 9), -1.0)
    pre[:, stage["inds_inside"]] = stage["labels_presample"]
    pre = pre.view(B, H, W, 9).permute(0, 3, 1, 2).reshape(B, -1)
    at = AnchorTargetLayer(16, [8, 16, 32], [0.5, 1, 2], rng=np.random.default_rng(5)).to(DEV)
    lab = at((torch.zeros(B, 18, H, W, device=DEV), gt.to(DEV), im_info.to(DEV), nb.to(DEV)))[0].cpu().view(B, -1)
    assert bool(((lab == pre) | (lab == -1)).all())
    for b in range(B):
        n_f, n_b = int((lab[b] == 1).sum()), int((lab[b] == 0).sum())
        assert n_f == min(128, int((pre[b] == 1).sum())) and n_f + n_b == min(256, n_f + int((pre[b] == 0).sum()))
    pt = ProposalTargetLayer(2, rng=np.random.default_rng(6))
    r, l, t, wi, wo = [x.cpu() for x in pt(rois.to(DEV), gt.to(DEV), nb.to(DEV))]
    assert r.shape == (B, 128, 5) and bool((l[:, 32:] == 0).all()) and bool((wi == wo).all())
    cand = torch.cat([rois[..., 1:], gt[..., :4]], 1)
    for b in range(B):
        assert bool((r[b, :, 0] == b).all())
        assert bool((r[b, :, None, 1:] == cand[b, None]).all(-1).any(-1).all()), "sampled roi is not a candidate"


def test_losses_match_reference_golden():
    """Five losses and their gradients w.r.t. the network outputs against the unmodified reference's autograd."""
    from ait_b200.targets import rpn_losses, rcnn_losses
    gold = load_golden("targets.pt")
    B = gold["shape"][0]
    rpn_cls_score, rpn_bbox_pred, score, bbox_pred = [t.to(DEV).requires_grad_() for t in _loss_inputs(gold)]
    a = [t.to(DEV) for t in gold["anchor_target"]]
    p = [t.to(DEV) for t in gold["proposal_target"]]
    l_rc, l_rb = rpn_losses(rpn_cls_score, rpn_bbox_pred, a)
    l_c, l_m, l_b = rcnn_losses(score, bbox_pred, p[1], p[2], p[3], p[4], B, margin=gold["margin"])
    for got, key in ((l_rc, "rpn_cls"), (l_rb, "rpn_box"), (l_c, "cls"), (l_m, "margin"), (l_b, "bbox")):
        assert torch.allclose(got.detach().cpu(), gold["losses"][key], rtol=2e-6, atol=0), (key, float(got), float(gold["losses"][key]))
    (l_rc + l_rb + l_c + l_m + l_b).backward()
    for t, key in ((rpn_cls_score, "rpn_cls_score"), (rpn_bbox_pred, "rpn_bbox_pred"), (score, "score"), (bbox_pred, "bbox_pred")):
        ref = gold["grads"][key]
        err = float((t.grad.cpu() - ref).abs().max() / ref.abs().max())
        assert err < 2e-5, (key, err)


def test_losses_weighted_sum_and_oracle_other_shapes():
    """Per-loss upstream gradients (a weighted sum) and a second shape (P = 100 rois per image, no foreground in one image)
    against the oracle's autograd."""
    from ait_b200.targets import rpn_losses, rcnn_losses
    from oracle import target_oracle as T
    g = torch.Generator().manual_seed(3)
    bs, P, A, H, W = 2, 100, 9, 12, 17
    score, bbox_pred = torch.randn(bs * P, 2, generator=g), torch.randn(bs * P, 4, generator=g)
    lab = (torch.rand(bs, P, generator=g) < 0.2).float()
    lab[1] = 0
    tgt = torch.randn(bs, P, 4, generator=g) * lab.unsqueeze(2)
    wi = lab.unsqueeze(2).expand(bs, P, 4).contiguous()
    rpn_s, rpn_b = torch.randn(bs, 2 * A, H, W, generator=g), torch.randn(bs, 4 * A, H, W, generator=g)
    rl = torch.randint(-1, 2, (bs, 1, A * H, W), generator=g).float()
    rt = torch.randn(bs, 4 * A, H, W, generator=g)
    rin = (rl.view(bs, A, 1, H, W) == 1).float().expand(bs, A, 4, H, W).reshape(bs, 4 * A, H, W).contiguous()
    rout = (rl.view(bs, A, 1, H, W) >= 0).float().expand(bs, A, 4, H, W).reshape(bs, 4 * A, H, W).contiguous() / 57.0
    wts = [0.7, 1.3, 2.0, 0.5, 1.1]

    def run(fn_rpn, fn_rcnn, dev):
        xs = [t.clone().to(dev).requires_grad_() for t in (rpn_s, rpn_b, score, bbox_pred)]
        l = list(fn_rpn(xs[0], xs[1], [t.to(dev) for t in (rl, rt, rin, rout)]))
        l += list(fn_rcnn(xs[2], xs[3], lab.view(-1).to(dev), tgt.to(dev), wi.to(dev), wi.to(dev), bs))
        sum(w * x for w, x in zip(wts, l)).backward()
        return [float(x.detach()) for x in l], [x.grad.cpu() for x in xs]

    ref_l, ref_g = run(lambda s, b, d: T.rpn_losses(s, b, *d), T.rcnn_losses, "cpu")
    got_l, got_g = run(lambda s, b, d: rpn_losses(s, b, d), rcnn_losses, DEV)
    for a_, b_ in zip(got_l, ref_l):
        assert abs(a_ - b_) <= 3e-6 * abs(b_)
    for a_, b_ in zip(got_g, ref_g):
        assert float((a_ - b_).abs().max()) <= 2e-5 * float(b_.abs().max())


def test_score_heads_training_step_matches_autograd():
    """feat / qfeat -> score heads -> the three detection losses -> backward: every gradient (both feature inputs, six
    parameter tensors) against torch autograd over the reference's nn.Linear modules + the loss oracle; and the adjoint
    of the 4x4 mean."""
    import torch.nn as nn
    from ait_b200.targets import score_heads, rcnn_losses, mean_pool_backward
    from oracle import target_oracle as T
    g = torch.Generator().manual_seed(11)
    bs, P = 3, 128
    G = bs * P
    feat, qfeat = torch.randn(G, 2048, generator=g).relu(), torch.randn(bs, 2048, generator=g).relu()
    lab = (torch.rand(bs, P, generator=g) < 0.25).float()
    tgt = torch.randn(bs, P, 4, generator=g) * lab.unsqueeze(2)
    wi = lab.unsqueeze(2).expand(bs, P, 4).contiguous()

    def make():
        torch.manual_seed(3)
        bb, cs = nn.Linear(2048, 4), nn.Sequential(nn.Linear(4096, 8), nn.Linear(8, 2))
        for m, std in ((bb, 0.01), (cs[0], 0.02), (cs[1], 0.3)):
            m.weight.data.normal_(0, std)
            m.bias.data.normal_(0, 0.1)
        return bb, cs

    def run(dev, ours):
        bb, cs = make()
        bb, cs = bb.to(dev), cs.to(dev)
        f, q = feat.clone().to(dev).requires_grad_(), qfeat.clone().to(dev).requires_grad_()
        if ours:
            score, bbox = score_heads(f, q, P, bb, cs)
            losses = rcnn_losses(score, bbox, lab.view(-1).to(dev), tgt.to(dev), wi.to(dev), wi.to(dev), bs)
        else:
            bbox = bb(f)
            score = cs(torch.cat((f.view(bs, P, -1), q.unsqueeze(1).repeat(1, P, 1)), 2).view(-1, 4096))
            losses = T.rcnn_losses(score, bbox, lab.view(-1), tgt, wi, wi, bs)
        (losses[0] + losses[1] + 2.0 * losses[2]).backward()
        grads = [f.grad, q.grad] + [p.grad for p in list(bb.parameters()) + list(cs.parameters())]
        return score.detach().cpu(), bbox.detach().cpu(), [x.cpu() for x in grads]

    s_ref, b_ref, g_ref = run("cpu", False)
    s, b, g_got = run(DEV, True)
    assert torch.allclose(s, s_ref, rtol=1e-4, atol=1e-5) and torch.allclose(b, b_ref, rtol=1e-4, atol=1e-5)
    names = ["d_feat", "d_qfeat", "bbox.weight", "bbox.bias", "cls0.weight", "cls0.bias", "cls1.weight", "cls1.bias"]
    for n, a_, r_ in zip(names, g_got, g_ref):
        err = float((a_ - r_).abs().max() / r_.abs().max())
        assert err < 1e-4, (n, err)
    d_top = mean_pool_backward(g_got[0].to(DEV)).cpu()
    assert torch.equal(d_top, (g_got[0] / 16).unsqueeze(1).expand(G, 16, 2048))


def test_target_layer_edge_cases():
    """An image without ground truth (anchor layer: all background, sampled to 256; proposal layer: the reference's
    ValueError), foreground-only and background-only images (the with-replacement branches), a single gt slot."""
    from ait_b200.targets import AnchorTargetLayer, ProposalTargetLayer
    from ait_b200.proposal import generate_anchors
    from oracle import target_oracle as T
    base = torch.from_numpy(generate_anchors(scales=np.array([8, 16, 32]), ratios=np.array([0.5, 1, 2]))).float()
    H, W = 19, 31
    im_info = torch.tensor([[300.0, 500.0, 1.0]] * 2)
    gt, nb = T.synth_gt_boxes(21, 2)
    gt[1] = 0                                        # image 1: no ground truth at all
    at = AnchorTargetLayer(16, [8, 16, 32], [0.5, 1, 2]).to(DEV)
    np.random.seed(5)
    ref = T.anchor_target(base, H, W, 16, gt, im_info)
    np.random.seed(5)
    out = at((torch.zeros(2, 18, H, W, device=DEV), gt.to(DEV), im_info.to(DEV), nb.to(DEV)))
    _check_anchor(out, ref)
    assert int((ref[0][1] == 1).sum()) == 0 and int((ref[0][1] == 0).sum()) == 256
    rois = T.synth_rois(22, 2, 100, gt)
    pt = ProposalTargetLayer(2)
    # effective training cfg (cfgs/res50.yml: BG_THRESH_LO 0.0): the rois of a gt-less image overlap nothing (0 >= 0.0), so they
    # are all background candidates -- same samples as the oracle; with config.py's never-used 0.1 the reference raises
    np.random.seed(8)
    ref = T.proposal_target(rois, gt)
    np.random.seed(8)
    _check_proposal(pt(rois.to(DEV), gt.to(DEV), nb.to(DEV)), ref)
    assert int((ref[1][1] > 0).sum()) == 0
    with pytest.raises(ValueError):
        ProposalTargetLayer(2, cfg=dict(BG_THRESH_LO=0.1))(rois.to(DEV), gt.to(DEV), nb.to(DEV))
    # foreground only (every roi sits on a gt box) / background only (rois overlap the gt a little, never >= 0.5)
    gt1 = torch.zeros(2, 1, 5)
    gt1[:, 0] = torch.tensor([100.0, 80.0, 260.0, 220.0, 1.0])
    r = torch.zeros(2, 60, 5)
    r[0, :, 1:] = gt1[0, 0, :4] + torch.randn(60, 4, generator=torch.Generator().manual_seed(1)) * 2.0
    r[1, :, 1:] = torch.tensor([180.0, 150.0, 420.0, 290.0]) + torch.randn(60, 4, generator=torch.Generator().manual_seed(2)) * 3.0
    r[1, :, 0] = 1
    stage = {}
    np.random.seed(6)
    ref = T.proposal_target(r, gt1, stage=stage)
    assert float(stage["max_overlaps"][0].min()) >= 0.5            # image 0: no background candidate
    np.random.seed(6)
    out = pt(r.to(DEV), gt1.to(DEV), None)
    _check_proposal(out, ref)
    assert bool((ref[1][0] == 1).all())


def test_device_rng_sampling():
    """rng="device": no host round trip.  Same invariants as the reference's sampling (counts, subset of the candidates,
    <= 32 distinct foreground rois first), deterministic per (seed, call), different across calls and seeds, and
    uniform: over many draws every candidate is kept about equally often."""
    from ait_b200.targets import AnchorTargetLayer, ProposalTargetLayer
    from ait_b200.proposal import generate_anchors
    from oracle import target_oracle as T
    B, H, W = 4, 38, 63
    gt, nb = T.synth_gt_boxes(9, B, im_h=600.0, im_w=1000.0, n_min=3, n_max=12)
    rois = T.synth_rois(10, B, 2000, gt, im_h=600.0, im_w=1000.0)
    im_info = torch.tensor([[600.0, 1000.0, 1.0]] * B)
    base = torch.from_numpy(generate_anchors(scales=np.array([8, 16, 32]), ratios=np.array([0.5, 1, 2]))).float()
    stage = {}
    T.anchor_target(base, H, W, 16, gt, im_info, stage=stage)
    pre = torch.full((B, H * W * 9), -1.0)
    pre[:, stage["inds_inside"]] = stage["labels_presample"]
    pre = pre.view(B, H, W, 9).permute(0, 3, 1, 2).reshape(B, -1)
    args = (torch.zeros(B, 18, H, W, device=DEV), gt.to(DEV), im_info.to(DEV), nb.to(DEV))

    def run_anchor(seed, calls=1):
        at = AnchorTargetLayer(16, [8, 16, 32], [0.5, 1, 2], rng="device", seed=seed).to(DEV)
        return [[t.cpu() for t in at(args)] for _ in range(calls)]

    a1, a2, a3 = run_anchor(1, 2), run_anchor(1, 2), run_anchor(2, 1)
    for x, y in zip(a1, a2):
        assert all(torch.equal(p, q) for p, q in zip(x, y)), "not deterministic for the same (seed, call)"
    assert not torch.equal(a1[0][0], a1[1][0]) and not torch.equal(a1[0][0], a3[0][0])
    lab = a1[0][0].view(B, -1)
    assert bool(((lab == pre) | (lab == -1)).all())
    for b in range(B):
        n_f, n_b = int((lab[b] == 1).sum()), int((lab[b] == 0).sum())
        assert n_f == min(128, int((pre[b] == 1).sum())) and n_f + n_b == min(256, n_f + int((pre[b] == 0).sum()))
    n_ex = int((lab[B - 1] >= 0).sum())
    assert torch.equal(a1[0][3] > 0, (lab >= 0).view(B, 9, 1, H, W).expand(B, 9, 4, H, W).reshape(B, 36, H, W))
    assert abs(float(a1[0][3].max()) - 1.0 / n_ex) < 1e-9
    # uniformity of the background draw of image 0: 60 calls, every candidate kept with frequency ~ k / n
    at = AnchorTargetLayer(16, [8, 16, 32], [0.5, 1, 2], rng="device", seed=7).to(DEV)
    cand = pre[0] == 0
    hits = torch.zeros(int(cand.sum()))
    calls = 60
    for _ in range(calls):
        hits += (at(args)[0].cpu().view(B, -1)[0][cand] == 0).float()
    n, k = hits.numel(), float(hits.sum()) / calls
    p = k / n
    z = (hits - calls * p) / (calls * p * (1 - p)) ** 0.5
    assert float(z.abs().max()) < 6.0 and abs(float(z.std()) - 1.0) < 0.15, (float(z.abs().max()), float(z.std()))

    pt = ProposalTargetLayer(2, rng="device", seed=3)
    r, l, t, wi, wo = [x.cpu() for x in pt(rois.to(DEV), gt.to(DEV), nb.to(DEV))]
    assert int(pt.last_bad_flag.item()) == 0
    assert r.shape == (B, 128, 5) and bool((l[:, 32:] == 0).all()) and bool((wi == wo).all())
    cand_boxes = torch.cat([rois[..., 1:], gt[..., :4]], 1)
    ref_stage = {}
    np.random.seed(0)
    T.proposal_target(rois, gt, stage=ref_stage)
    for b in range(B):
        assert bool((r[b, :, 0] == b).all())
        match = (r[b, :, None, 1:] == cand_boxes[b, None]).all(-1)
        assert bool(match.any(-1).all()), "sampled roi is not a candidate"
        idx = match.float().argmax(-1)
        mo = ref_stage["max_overlaps"][b][idx]
        n_fg = min(32, int((ref_stage["max_overlaps"][b] >= 0.5).sum()))
        assert bool((mo[:n_fg] >= 0.5).all()) and bool(((mo[n_fg:] < 0.5) & (mo[n_fg:] >= 0.0)).all())
        assert len(set(idx[:n_fg].tolist())) == n_fg, "foreground rois must be drawn without replacement"
        assert bool((l[b, :n_fg] == 1).all())
    # an image without any candidate: flagged on the device instead of raised (no synchronisation in this mode)
    gt0 = gt.clone()
    gt0[1] = 0
    rois0 = rois.clone()
    rois0[1, :, 1:] = 0
    pt(rois0.to(DEV), gt0.to(DEV), nb.to(DEV))
    assert int(pt.last_bad_flag.item()) == 1
