"""Per-stage error of the GPU head against the CPU oracle (fp32) and the oracle in fp64 (run on the GPU box)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402

from conftest import golden_head, head_inputs  # noqa: E402
from oracle import head_oracle  # noqa: E402


def err(a, b):
    a, b = a.double(), b.double()
    d = (a - b).abs()
    return "max|d|/max|ref| %.2e  rms(d)/rms(ref) %.2e" % (float(d.max() / b.abs().max()),
                                                          float(d.pow(2).mean().sqrt() / b.pow(2).mean().sqrt()))


def main():
    B, P = 2, 4
    non_img, non_qry, rois = head_inputs(B, P)
    for dtype in ("fp32", "tf32", "bf16"):
        for round_acts in ((False, True) if dtype == "tf32" else (False,)):
            head, g = golden_head(compute_dtype=dtype)
            sd = head.state_dict()
            with torch.no_grad():
                r32 = head_oracle.head_forward(sd, non_img, non_qry, rois)
                r64 = head_oracle.head_forward(sd, non_img, non_qry, rois, dtype=torch.float64)
            head = head.cuda()
            from ait_b200 import packing
            head._engine = packing.HeadEngine(transformer=head.transformer, sk=head.sk, top=head.RCNN_top,
                                              cls_score=head.RCNN_cls_score, bbox_pred=head.RCNN_bbox_pred,
                                              dtype=dtype, round_acts=round_acts)
            cls, bbox, t = head(non_img.cuda(), non_qry.cuda(), rois.cuda(), taps=True)
            bp = B * P
            mine = dict(pooled=t["pooled"].float().cpu().permute(0, 2, 1).reshape(bp, 1024, 7, 7),
                        enc_out=t["enc_out"].float().cpu(),
                        ait_out=t["ait_out"].float().cpu().permute(0, 2, 1).reshape(bp, 1024, 8, 8),
                        sk_out=t["sk_out"].float().cpu().permute(0, 2, 1).reshape(bp, 1024, 8, 8),
                        feat=t["feat"].cpu(), qfeat=t["qfeat"].cpu(), bbox_pred=bbox.cpu(), cls_prob=cls.cpu())
            print("==== dtype", dtype, "round_acts", round_acts)
            for k in ("pooled", "enc_out", "ait_out", "sk_out", "feat", "qfeat", "bbox_pred", "cls_prob"):
                a, b32, b64 = mine[k], r32[k], r64[k]
                if k == "enc_out":
                    a, b32, b64 = a[:, :49], b32[:, :49], b64[:, :49]
                print("%-9s gpu-vs-fp64: %s | cpu32-vs-fp64: %s" % (k, err(a, b64), err(b32, b64)))
            print("cls_prob gpu  ", [round(x, 4) for x in cls.flatten().tolist()])
            print("cls_prob fp64 ", [round(x, 4) for x in r64["cls_prob"].flatten().tolist()])
            print("cls_prob cpu32", [round(x, 4) for x in r32["cls_prob"].flatten().tolist()])
            # conditioning of the calibrated score layer: logit change for a 1e-3 relative feature perturbation
            f = r64["feat"]
            gen = torch.Generator().manual_seed(0)
            pert = f * (1 + 1e-3 * torch.randn(f.shape, generator=gen, dtype=torch.float64))
            w = {k: v.double() for k, v in sd.items()}
            def logits(pf):
                st = torch.cat([pf.view(B, P, -1), r64["qfeat"].unsqueeze(1).repeat(1, P, 1)], 2).view(-1, 4096)
                h = torch.nn.functional.linear(st, w["RCNN_cls_score.0.weight"], w["RCNN_cls_score.0.bias"])
                return torch.nn.functional.linear(h, w["RCNN_cls_score.1.weight"], w["RCNN_cls_score.1.bias"])
            print("logit shift for 1e-3 relative feature noise:", float((logits(pert) - logits(f)).abs().max()))


if __name__ == "__main__":
    main()
