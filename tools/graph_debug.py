"""Which part of the whole-head training step breaks CUDA-graph capture?  python tools/graph_debug.py <stage> [mode]
stages: roi | ait | sk | top | heads | fwd | fwdbwd"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ait_b200 import sk_train, synth, targets, top_train  # noqa: E402

stage = sys.argv[1]
mode = sys.argv[2] if len(sys.argv) > 2 else "global"
dev = "cuda:0"
B, P = int(os.environ.get("GB", 2)), int(os.environ.get("GP", 8))
head = synth.make_head(seed=0, calibrated=True, randomize_bn=True)
from ait_b200.system.Models import set_dropout
set_dropout(head, 0.0, 0.0)
head = head.to(dev).train()
maps = torch.stack([synth.c4_map(u) for u in range(B)]).to(dev).requires_grad_()
qrys = torch.stack([synth.query_feat(u) for u in range(B)]).to(dev).requires_grad_()
rois = torch.stack([synth.random_rois(u, P, batch_index=u) for u in range(B)]).to(dev)
label = (torch.arange(B * P) % 3 == 0).long().to(dev)
tgt = torch.zeros(B * P, 4, device=dev)
inw = (label > 0).float().view(-1, 1).expand(-1, 4).contiguous()
xp = torch.randn(B * P, 1024, 7, 7, device=dev).relu().requires_grad_()
x8 = torch.randn(B * P, 1024, 8, 8, device=dev).relu().requires_grad_()
f1 = torch.randn(B * P, 2048, device=dev).requires_grad_()
f2 = torch.randn(B, 2048, device=dev).requires_grad_()


def run():
    head.zero_grad(set_to_none=True)
    for t in (maps, qrys, xp, x8, f1, f2):
        t.grad = None
    if stage == "roi":
        out = head.RCNN_roi_align(maps, rois.reshape(-1, 5)); out.sum().backward()
    elif stage == "ait":
        out = head.transformer(xp, qrys); out.sum().backward()
    elif stage == "aitfwd":
        with torch.no_grad():
            pass
        out = head.transformer(xp, qrys)
    elif stage == "sk":
        a, b = sk_train.sknet_train(head.sk, x8, qrys); (a.sum() + b.sum()).backward()
    elif stage == "top":
        out = top_train.head_to_tail_train(head.RCNN_top, x8); out.sum().backward()
    elif stage == "heads":
        s, bb = targets.score_heads(f1, f2, P, head.RCNN_bbox_pred, head.RCNN_cls_score)
        sum(targets.rcnn_losses(s, bb, label, tgt, inw, inw, B)).backward()
    elif stage == "fwd":
        head.forward_train(maps, qrys, rois)
    else:
        sum(head.training_losses(maps, qrys, rois, label, tgt, inw, inw)).backward()


side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    run()
    run()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
cg = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(cg, capture_error_mode=mode):
        run()
    cg.replay()
    torch.cuda.synchronize()
    print(stage, mode, "CAPTURE OK")
except Exception as e:
    print(stage, mode, "FAILED:", type(e).__name__, str(e)[:200].replace("\n", " "))
