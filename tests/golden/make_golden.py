"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_import.py) on seeded synthetic inputs.

    python tests/golden/make_golden.py          (build container only: needs /root/reference)

The reference has no tests or fixtures of its own, so these files ARE the
pins: every tensor below is produced by the reference's own modules / CPU kernels with the weights
of `ait_b200.synth.make_head(seed=0, calibrated=True, randomize_bn=True)` loaded through
`load_state_dict(strict=True)` -- which also proves state_dict compatibility.
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from ait_b200 import synth  # noqa: E402
from oracle import ref_import  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    torch.set_num_threads(8)
    ref_import.install()
    from model.roi_layers import ROIAlign, nms
    from model.rpn.generate_anchors import generate_anchors

    head = synth.make_head(seed=0, calibrated=True, randomize_bn=True)
    sd = head.state_dict()

    # ---------------- reference modules with OUR weights (strict)
    T = ref_import.ref_transformer().eval()
    T.load_state_dict(head.transformer.state_dict(), strict=True)
    SK = ref_import.ref_sknet().eval()
    SK.load_state_dict(head.sk.state_dict(), strict=True)
    L4 = ref_import.ref_layer4().eval()
    L4.load_state_dict(head.RCNN_top.state_dict(), strict=True)
    cls_score = torch.nn.Sequential(torch.nn.Linear(4096, 8), torch.nn.Linear(8, 2))
    cls_score.load_state_dict(head.RCNN_cls_score.state_dict())
    bbox_pred = torch.nn.Linear(2048, 4)
    bbox_pred.load_state_dict(head.RCNN_bbox_pred.state_dict())

    keys = {
        "transformer": {k: list(v.shape) for k, v in T.state_dict().items()},
        "sk": {k: list(v.shape) for k, v in SK.state_dict().items()},
        "RCNN_top": {k: list(v.shape) for k, v in L4.state_dict().items()},
    }
    with open(os.path.join(OUT, "state_dict_keys.json"), "w") as f:
        json.dump(keys, f, indent=1, sort_keys=True)

    # ---------------- head: B=2 units x P=4 proposals (faster_rcnn_coatt_transformer_sk.py:273-337)
    B, P = 2, 4
    non_img = torch.stack([synth.c4_map(u) for u in range(B)])
    non_qry = torch.stack([synth.query_feat(u) for u in range(B)])
    rois = torch.stack([synth.random_rois(u, P, batch_index=u) for u in range(B)])
    with torch.no_grad():
        ra = ROIAlign((7, 7), 1.0 / 16.0, 0)
        pooled = ra(non_img, rois.view(-1, 5))
        ait = T(x_props=pooled, x_query=non_qry)
        sp, sq = SK(x_props=ait, x_query=non_qry)
        pf = L4(sp).mean(3).mean(2)
        qf = L4(sq).mean(3).mean(2)
        bbox = bbox_pred(pf)
        stack = torch.cat((pf.view(B, P, -1), qf.unsqueeze(1).repeat(1, P, 1)), dim=2).view(-1, 4096)
        # Spread-calibrated RCNN_cls_score (the stock init gives cls_prob = 0.0057 +- 1e-6,
        # which would make the 1e-3 absolute gate vacuous).  The first layer reads the three principal
        # directions of these 8 feature vectors (what a trained similarity head does: weights aligned with
        # the discriminative directions), each normalised to unit spread; logits then differ by O(1) while
        # staying well conditioned (a random direction would amplify feature noise by |x| / |delta x|).
        Xc = stack - stack.mean(dim=0, keepdim=True)
        _, S, Vt = torch.linalg.svd(Xc, full_matrices=False)
        A = torch.zeros(8, 4096)
        for j in range(3):
            A[j] = Vt[j] / (S[j] / 8 ** 0.5)
        b1 = -(stack @ A.t()).mean(dim=0)
        wv = torch.tensor([0.9, -0.7, 0.5, 0, 0, 0, 0, 0])
        W2 = torch.stack([-wv, wv])
        cls_score[0].weight.copy_(A)
        cls_score[0].bias.copy_(b1)
        cls_score[1].weight.copy_(W2)
        cls_score[1].bias.zero_()
        score = cls_score(stack)
        prob = torch.nn.functional.softmax(score, 1)[:, 1]
    torch.save(dict(B=B, P=P, rois=rois,
                    pooled_s=pooled[:, ::16].clone(), ait_s=ait[:, ::16].clone(), sk_s=sp[:, ::16].clone(),
                    feat=pf, qfeat=qf, score=score, cls_score_state=cls_score.state_dict(), cls_prob=prob.view(B, P, 1), bbox_pred=bbox.view(B, P, 4),
                    pooled_sha=sha(pooled.numpy()), ait_sha=sha(ait.numpy())),
               os.path.join(OUT, "head_b2p4.pt"))
    print("head: cls_prob", prob.tolist())

    # ---------------- AIT module alone, adaptive_image_transformer.py-style inputs (torch.rand), bs=2 x 3
    g = torch.Generator().manual_seed(7)
    xp = torch.rand(6, 1024, 7, 7, generator=g)
    xq = torch.rand(2, 1024, 8, 8, generator=g)
    with torch.no_grad():
        out = T(x_props=xp, x_query=xq)
    torch.save(dict(seed=7, out_s=out[:, ::8].clone(), out_sha=sha(out.numpy())), os.path.join(OUT, "ait_rand.pt"))

    # ---------------- ROIAlign: small map, edge-case rois (ROIAlign_cpu.cpp:113-219)
    g = torch.Generator().manual_seed(11)
    feat = torch.randn(2, 8, 38, 63, generator=g)
    rr = torch.tensor([
        [0, 10.0, 20.0, 300.0, 400.0],
        [1, 0.0, 0.0, 999.0, 599.0],          # whole image: grid 9 x 6
        [0, 500.3, 100.7, 507.9, 104.2],      # smaller than one cell -> forced 1x1
        [1, 990.0, 590.0, 999.0, 599.0],      # bottom-right corner: clamping
        [0, -40.0, -30.0, 60.0, 50.0],        # partly outside (y < -1 samples contribute 0)
        [1, 900.0, 500.0, 1200.0, 800.0],     # extends past the map
        [0, 100.0, 100.0, 100.0, 100.0],      # degenerate
        [1, 333.3, 77.7, 666.6, 555.5],
    ])
    ref_out = ROIAlign((7, 7), 1.0 / 16.0, 0)(feat, rr)
    ref_out2 = ROIAlign((7, 7), 1.0 / 16.0, 2)(feat, rr)
    torch.save(dict(seed=11, rois=rr, out=ref_out, out_sr2=ref_out2), os.path.join(OUT, "roi_align_small.pt"))

    # ---------------- NMS at the RPN site: unit 0, VOC anchors, top 6000, thr 0.7 (proposal_layer.py:129-157)
    boxes, scores = synth.rpn_outputs(0)
    order = torch.sort(scores, 0, True)[1][:6000]
    b6, s6 = boxes[order], scores[order]
    keep = nms(b6, s6, 0.7).long()          # reference CPU kernel (>=); inputs have no exact ties at 0.7
    keep_all = nms(boxes, scores, 0.7).long()
    rois0 = torch.zeros(300, 5)
    k300 = keep[:300]
    rois0[: len(k300), 1:] = b6[k300]
    # anchors
    anchors = {"voc": generate_anchors(scales=np.array([8, 16, 32])).tolist(),
               "coco": generate_anchors(scales=np.array([4, 8, 16, 32])).tolist()}
    torch.save(dict(unit=0, n_pre=6000, thr=0.7, n_keep=int(keep.numel()), keep_first300=k300.clone(),
                    keep_sha=sha(keep.numpy()), rois=rois0, n_keep_all=int(keep_all.numel()),
                    keep_all_sha=sha(keep_all.numpy()), keep_all_first64=keep_all[:64].clone(), anchors=anchors),
               os.path.join(OUT, "nms_rpn_unit0.pt"))
    print("nms: kept", keep.numel(), "of 6000;", keep_all.numel(), "of", boxes.shape[0])
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
