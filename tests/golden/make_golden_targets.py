"""Golden outputs of the reference's training-only samplers and losses (row f4).

    python tests/golden/make_golden_targets.py     (build container only: needs /root/reference)

Runs the UNMODIFIED `_AnchorTargetLayer`, `_ProposalTargetLayer`, `_smooth_l1_loss`, the RPN loss lines of
`_RPN.forward` and the detection-loss lines of `_fasterRCNN.forward` (through F.cross_entropy /
torch.nn.MarginRankingLoss exactly as the reference calls them) after `np.random.seed`, and writes
tests/golden/targets.pt.  Inputs are regenerated from the seeds by the tests (oracle/target_oracle.py).
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import ref_import, target_oracle as T  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
B, A, H, W, R = 3, 9, 19, 31, 400


def loss_inputs(seed=47):
    g = torch.Generator().manual_seed(seed)
    rpn_cls_score = torch.randn(B, 2 * A, H, W, generator=g)
    rpn_bbox_pred = 0.4 * torch.randn(B, 4 * A, H, W, generator=g)
    score = torch.randn(B * 128, 2, generator=g)
    bbox_pred = 0.8 * torch.randn(B * 128, 4, generator=g)
    return rpn_cls_score, rpn_bbox_pred, score, bbox_pred


def main():
    ref_import.install()
    from model.rpn.anchor_target_layer import _AnchorTargetLayer
    from model.rpn.proposal_target_layer_cascade import _ProposalTargetLayer
    from model.utils.net_utils import _smooth_l1_loss
    from model.utils.config import cfg, cfg_from_file
    # the EFFECTIVE training configuration: trainval_net_voc.py:206-209 always merges one of cfgs/res50*.yml / res101.yml
    # over the config.py defaults, and every one of them sets TRAIN.BG_THRESH_LO: 0.0 (config.py's 0.1 is never used)
    cfg_from_file(os.path.join(ref_import.REF_ROOT, "cfgs", "res50.yml"))
    assert cfg.TRAIN.BG_THRESH_LO == 0.0 and cfg.TRAIN.BATCH_SIZE == 128
    gt, nb = T.synth_gt_boxes(41, B)
    im_info = torch.tensor([[300.0, 500.0, 1.0]] * B)
    rois = T.synth_rois(43, B, R, gt)
    at = _AnchorTargetLayer(16, [8, 16, 32], [0.5, 1, 2])
    np.random.seed(7)
    a_out = at((torch.zeros(B, 2 * A, H, W), gt, im_info, nb))
    pt = _ProposalTargetLayer(2)
    np.random.seed(11)
    p_out = pt(rois, gt, nb)

    rpn_cls_score, rpn_bbox_pred, score, bbox_pred = [t.requires_grad_() for t in loss_inputs()]
    # rpn.py:99-126
    rpn_cls_score_reshape = rpn_cls_score.view(B, 2, -1, W)
    s = rpn_cls_score_reshape.permute(0, 2, 3, 1).contiguous().view(B, -1, 2)
    rpn_label = a_out[0].view(B, -1)
    keep = rpn_label.view(-1).ne(-1).nonzero().view(-1)
    rpn_loss_cls = F.cross_entropy(torch.index_select(s.view(-1, 2), 0, keep),
                                   torch.index_select(rpn_label.view(-1), 0, keep).long())
    rpn_loss_box = _smooth_l1_loss(rpn_bbox_pred, a_out[1], a_out[2], a_out[3], sigma=3, dim=[1, 2, 3])
    # faster_rcnn_coatt_transformer_sk.py:334-361
    rois_label = p_out[1].view(-1).long()
    score_prob = F.softmax(score, 1)[:, 1]
    score_label = rois_label.view(B, -1).float()
    gt_map = torch.abs(score_label.unsqueeze(1) - score_label.unsqueeze(-1))
    sp = score_prob.view(B, -1)
    pr_map = torch.abs(sp.unsqueeze(1) - sp.unsqueeze(-1))
    target = -((gt_map - 1) ** 2) + gt_map
    loss_cls = F.cross_entropy(score, rois_label)
    margin_loss = 3 * torch.nn.MarginRankingLoss(margin=cfg.TRAIN.MARGIN)(pr_map, gt_map, target)
    loss_bbox = _smooth_l1_loss(bbox_pred, p_out[2].view(-1, 4), p_out[3].view(-1, 4), p_out[4].view(-1, 4))
    total = rpn_loss_cls + rpn_loss_box + loss_cls + margin_loss + loss_bbox
    total.backward()
    torch.save(dict(cfg_file="cfgs/res50.yml", bg_thresh_lo=float(cfg.TRAIN.BG_THRESH_LO), seeds=dict(gt=41, rois=43, anchor_np=7, proposal_np=11, loss=47), shape=(B, A, H, W, R),
                    margin=float(cfg.TRAIN.MARGIN), anchors=at._anchors.clone(),
                    anchor_target=[t.clone() for t in a_out], proposal_target=[t.clone() for t in p_out],
                    losses=dict(rpn_cls=rpn_loss_cls.detach(), rpn_box=rpn_loss_box.detach(), cls=loss_cls.detach(),
                                margin=margin_loss.detach(), bbox=loss_bbox.detach()),
                    grads=dict(rpn_cls_score=rpn_cls_score.grad.clone(), rpn_bbox_pred=rpn_bbox_pred.grad.clone(),
                               score=score.grad.clone(), bbox_pred=bbox_pred.grad.clone())),
               os.path.join(OUT, "targets.pt"))
    print("wrote targets.pt; anchor labels (-1,0,1):", [(a_out[0] == v).sum().item() for v in (-1, 0, 1)],
          "fg rois per image:", (p_out[1] > 0).sum(1).tolist(),
          "losses:", [float(x) for x in (rpn_loss_cls, rpn_loss_box, loss_cls, margin_loss, loss_bbox)])


if __name__ == "__main__":
    main()
