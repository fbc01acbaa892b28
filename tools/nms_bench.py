"""NMS / top-n latency on the benchmark inputs (CUDA events, hot):  python tools/nms_bench.py
  proposal mode: top-6000 of 21 546 + kept-list NMS 0.7 -> 300, 8 images per launch
  mask mode    : nms(dets, scores, thr) drop-in on one image's top-6000 / all 21 546 boxes (bitmask + on-device scan)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ait_b200 import ops, synth  # noqa: E402
from ait_b200.proposal import propose_rois  # noqa: E402
from ait_b200.roi_layers import nms  # noqa: E402

dev = "cuda:0"
B = 8
rpn = [synth.rpn_outputs(u) for u in range(B)]
boxes, scores = torch.stack([r[0] for r in rpn]).to(dev), torch.stack([r[1] for r in rpn]).to(dev)


def t(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return round(s.elapsed_time(e) / n * 1e3, 1)


res = {}
res["propose_rois_8img_us"] = t(lambda: propose_rois(boxes, scores, 6000, 300, 0.7))
order = ops.topk_desc(scores, 6000)
res["topk_6000_8img_us"] = t(lambda: ops.topk_desc(scores, 6000))
res["nms_lazy_8img_us"] = t(lambda: ops.nms_batched(boxes, order, 0.7, 300, mode=0, want_rois=True))
res["nms_lazy_2000_8img_us"] = t(lambda: ops.nms_batched(boxes, order, 0.7, 2000, mode=0, want_rois=True))
o0 = order[0]
d0, s0 = boxes[0][o0].contiguous(), scores[0][o0].contiguous()
res["nms_dropin_6000_us"] = t(lambda: nms(d0, s0, 0.7))
res["nms_dropin_6000_kept"] = int(nms(d0, s0, 0.7).numel())
ord1 = torch.arange(6000, device=dev).view(1, -1)
res["nms_mask_scan_only_6000_us"] = t(lambda: ops.nms_batched(d0.view(1, -1, 4), ord1, 0.7, 6000, mode=1))
res["nms_dropin_21546_us"] = t(lambda: nms(boxes[0], scores[0], 0.7), 5)
res["nms_dropin_21546_kept"] = int(nms(boxes[0], scores[0], 0.7).numel())
print(json.dumps(res))
