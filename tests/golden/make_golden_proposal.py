"""Golden output of the reference's whole proposal layer (row f1) on seeded synthetic RPN outputs.

    python tests/golden/make_golden_proposal.py     (build container only: needs /root/reference)

Runs the UNMODIFIED `_ProposalLayer` (lib/model/rpn/proposal_layer.py) with the reference's own CPU nms
and writes tests/golden/proposal_layer.pt (inputs are regenerated from the seed by the tests).
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import ref_import  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def inputs(seed=21, B=2, A=9, H=19, W=31):
    g = torch.Generator().manual_seed(seed)
    cls_prob = torch.rand(B, 2 * A, H, W, generator=g)
    bbox_pred = 0.3 * torch.randn(B, 4 * A, H, W, generator=g)
    im_info = torch.tensor([[300.0, 500.0, 1.5], [280.0, 480.0, 0.8]])[:B]
    return cls_prob, bbox_pred, im_info


def main():
    ref_import.install()
    from model.rpn.proposal_layer import _ProposalLayer
    from model.utils.config import cfg
    cls_prob, bbox_pred, im_info = inputs()
    layer = _ProposalLayer(16, [8, 16, 32], [0.5, 1, 2])
    cfg.TEST.RPN_PRE_NMS_TOP_N, cfg.TEST.RPN_POST_NMS_TOP_N, cfg.TEST.RPN_NMS_THRESH = 3000, 100, 0.7
    rois = layer((cls_prob, bbox_pred, im_info, "TEST"))
    torch.save(dict(seed=21, pre=3000, post=100, thr=0.7, rois=rois.clone(), anchors=layer._anchors.clone()),
               os.path.join(OUT, "proposal_layer.pt"))
    print("wrote proposal_layer.pt", tuple(rois.shape), "non-zero rows:", int((rois[..., 1:].abs().sum(-1) > 0).sum()))


if __name__ == "__main__":
    main()
