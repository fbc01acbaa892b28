"""Golden outputs of the reference's RPN head (`_RPN`, lib/model/rpn/rpn.py, eval mode) on a seeded map.

    python tests/golden/make_golden_rpn.py     (build container only: needs /root/reference)
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import ref_import  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def inputs(seed=23, B=2, H=19, W=31):
    g = torch.Generator().manual_seed(seed)
    base_feat = torch.randn(B, 1024, H, W, generator=g).relu()
    im_info = torch.tensor([[300.0, 500.0, 1.5], [280.0, 480.0, 0.8]])[:B]
    return base_feat, im_info


def weights(seed=23):
    """state_dict of the RPN head with enough spread for non-degenerate scores / boxes."""
    g = torch.Generator().manual_seed(seed + 1)
    return {"RPN_Conv.weight": torch.randn(512, 1024, 3, 3, generator=g) * 0.01,
            "RPN_Conv.bias": torch.randn(512, generator=g) * 0.1,
            "RPN_cls_score.weight": torch.randn(18, 512, 1, 1, generator=g) * 0.05,
            "RPN_cls_score.bias": torch.randn(18, generator=g) * 0.1,
            "RPN_bbox_pred.weight": torch.randn(36, 512, 1, 1, generator=g) * 0.01,
            "RPN_bbox_pred.bias": torch.randn(36, generator=g) * 0.05}


def main():
    torch.set_num_threads(8)
    ref_import.install()
    from model.rpn.rpn import _RPN
    from model.utils.config import cfg
    cfg.TEST.RPN_PRE_NMS_TOP_N, cfg.TEST.RPN_POST_NMS_TOP_N, cfg.TEST.RPN_NMS_THRESH = 3000, 100, 0.7
    rpn = _RPN(1024).eval()
    missing = rpn.load_state_dict(weights(), strict=False)
    assert not missing.unexpected_keys, missing
    base_feat, im_info = inputs()
    captured = {}
    orig = rpn.RPN_proposal.forward

    def spy(inp):
        captured["cls_prob"], captured["bbox_pred"] = inp[0].clone(), inp[1].clone()
        return orig(inp)

    rpn.RPN_proposal.forward = spy
    with torch.no_grad():
        rois, _, _ = rpn(base_feat, im_info, None, None)
    torch.save(dict(seed=23, pre=3000, post=100, thr=0.7, rois=rois.clone(), cls_prob=captured["cls_prob"],
                    bbox_pred_s=captured["bbox_pred"][:, ::3].clone(), anchors=rpn.RPN_proposal._anchors.clone()),
               os.path.join(OUT, "rpn_head.pt"))
    print("wrote rpn_head.pt", tuple(rois.shape), "fg prob range", float(captured["cls_prob"][:, 9:].min()),
          float(captured["cls_prob"][:, 9:].max()))


if __name__ == "__main__":
    main()
