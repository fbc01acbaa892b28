"""Golden GRADIENTS of the whole detection-head training step (BASELINE config 4) from the UNMODIFIED reference modules run on
CPU fp32 with torch autograd: model.roi_layers.ROIAlign (forward only -- the reference has no CPU ROIAlign backward,
ROIAlign.h:44 -- so the pooled features are the differentiated leaf) -> model.system.Models.Transformer (.train(), every
nn.Dropout p = 0) -> model.modules...SKNet -> RCNN_top = layer4 (frozen BatchNorm, eval) + mean -> nn.Linear heads -> the
detection-loss lines of `_fasterRCNN.forward` (faster_rcnn_coatt_transformer_sk.py:340-361: F.cross_entropy,
3 * MarginRankingLoss(cfg.TRAIN.MARGIN), _smooth_l1_loss).

    python tests/golden/make_golden_head_grad.py     (build container only: needs /root/reference)

Writes tests/golden/head_grad.pt: the three losses, the gradient of the pooled features and of the query feature (strided
samples + norms) and, for every parameter that receives a gradient, its L2 norm plus a strided sample.
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from ait_b200 import synth  # noqa: E402
from oracle import ref_import  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SAMPLE = 1024
B, P = 2, 4


def inputs():
    g = torch.Generator().manual_seed(71)
    maps = torch.stack([synth.c4_map(u) for u in range(B)])
    qrys = torch.stack([synth.query_feat(u) for u in range(B)])
    rois = torch.stack([synth.random_rois(u, P, batch_index=u) for u in range(B)])
    label = torch.tensor([[1, 0, 0, 1], [0, 0, 1, 0]]).view(-1)
    tgt = 0.3 * torch.randn(B * P, 4, generator=g)
    inw = (label > 0).float().view(-1, 1).expand(-1, 4).contiguous()
    return maps, qrys, rois, label, tgt, inw, inw.clone()


def sample(t):
    f = t.reshape(-1)
    if f.numel() <= SAMPLE:
        return f.clone(), 1
    stride = f.numel() // SAMPLE
    return f[::stride][:SAMPLE].clone(), stride


def main():
    torch.set_num_threads(8)
    ref_import.install()
    from model.roi_layers import ROIAlign
    from model.utils.config import cfg
    from model.utils.net_utils import _smooth_l1_loss

    head = synth.make_head(seed=0, calibrated=True, randomize_bn=True)
    T = ref_import.ref_transformer(dropout=0.0).train()
    for mod in T.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    T.load_state_dict(head.transformer.state_dict(), strict=True)
    SK = ref_import.ref_sknet().train()
    SK.load_state_dict(head.sk.state_dict(), strict=True)
    L4 = ref_import.ref_layer4().eval()                 # frozen BatchNorm: set_bn_fix + .eval() in the reference's train()
    L4.load_state_dict(head.RCNN_top.state_dict(), strict=True)
    for n, p in L4.named_parameters():
        p.requires_grad_(p.dim() == 4)                  # BatchNorm parameters do not train (resnet...:429-435)
    cls_score = torch.nn.Sequential(torch.nn.Linear(4096, 8), torch.nn.Linear(8, 2))
    cls_score.load_state_dict(head.RCNN_cls_score.state_dict())
    bbox_pred_l = torch.nn.Linear(2048, 4)
    bbox_pred_l.load_state_dict(head.RCNN_bbox_pred.state_dict())

    maps, qrys, rois, label, tgt, inw, outw = inputs()
    with torch.no_grad():
        pooled = ROIAlign((7, 7), 1.0 / 16.0, 0)(maps, rois.view(-1, 5))
    pooled.requires_grad_()
    qrys.requires_grad_()
    # faster_rcnn_coatt_transformer_sk.py:289-335
    props = T(x_props=pooled, x_query=qrys)
    props, query = SK(x_props=props, x_query=qrys)
    pf = L4(props).mean(3).mean(2)
    qf = L4(query).mean(3).mean(2)
    bbox_pred = bbox_pred_l(pf)
    stack = torch.cat((pf.view(B, P, -1), qf.unsqueeze(1).repeat(1, P, 1)), dim=2).view(-1, 4096)
    score = cls_score(stack)
    score_prob = F.softmax(score, 1)[:, 1]
    # :340-361
    score_label = label.view(B, -1).float()
    gt_map = torch.abs(score_label.unsqueeze(1) - score_label.unsqueeze(-1))
    pr = score_prob.view(B, -1)
    pr_map = torch.abs(pr.unsqueeze(1) - pr.unsqueeze(-1))
    target = -((gt_map - 1) ** 2) + gt_map
    loss_cls = F.cross_entropy(score, label)
    margin_loss = 3 * torch.nn.MarginRankingLoss(margin=cfg.TRAIN.MARGIN)(pr_map, gt_map, target)
    loss_bbox = _smooth_l1_loss(bbox_pred, tgt, inw, outw)
    (loss_cls + margin_loss + loss_bbox).backward()

    params = {}
    for prefix, mod in (("transformer.", T), ("sk.", SK), ("RCNN_top.", L4), ("RCNN_cls_score.", cls_score),
                        ("RCNN_bbox_pred.", bbox_pred_l)):
        for name, p in mod.named_parameters():
            if p.grad is None:
                continue
            s, stride = sample(p.grad)
            params[prefix + name] = dict(norm=float(p.grad.double().norm()), sample=s, stride=stride)
    gp, gps = sample(pooled.grad)
    gq, gqs = sample(qrys.grad)
    torch.save(dict(seed=71, B=B, P=P, losses=[float(loss_cls), float(margin_loss), float(loss_bbox)],
                    score=score.detach().clone(), bbox_pred=bbox_pred.detach().clone(),
                    grad_pooled=dict(norm=float(pooled.grad.double().norm()), sample=gp, stride=gps),
                    grad_query=dict(norm=float(qrys.grad.double().norm()), sample=gq, stride=gqs), params=params),
               os.path.join(OUT, "head_grad.pt"))
    print("wrote head_grad.pt: losses", [float(loss_cls), float(margin_loss), float(loss_bbox)], len(params),
          "parameter gradients; |grad_pooled| = %.5f" % float(pooled.grad.norm()))


if __name__ == "__main__":
    main()
