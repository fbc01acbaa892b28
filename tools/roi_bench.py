"""ROIAlign (token-major, benchmark rois: 8 units x 300 proposals from the proposal layer) hot, CUDA events.
    python tools/roi_bench.py            per-roi kernel;   AITB_ROI_SLAB=1 python tools/roi_bench.py   (roi, slab) kernel"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ait_b200 import _lib as L, ops, synth  # noqa: E402
from ait_b200.proposal import propose_rois  # noqa: E402

dev = "cuda:0"
B, P = 8, 300
maps = torch.stack([synth.c4_map(u) for u in range(B)]).to(dev)
rpn = [synth.rpn_outputs(u) for u in range(B)]
boxes, scores = torch.stack([r[0] for r in rpn]).to(dev), torch.stack([r[1] for r in rpn]).to(dev)
rois, _ = propose_rois(boxes, scores)
lib = L.load()


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
    return s.elapsed_time(e) / n * 1e3


res = {"kernel": "(roi, 128-channel slab) CTAs"}
nhwc = ops.transpose_cs(maps.reshape(B, 1024, -1), True, out_dtype=torch.float32).view(B, 38, 63, 1024)
pooled = torch.empty(B * P, 49, 2048, device=dev, dtype=torch.bfloat16)
res["fp32_split_us"] = round(t(lambda: L.check(lib.aitb_roi_align_forward(
    L.ptr(nhwc), L.ptr(rois.view(-1, 5)), B, 1024, 38, 63, B * P, 1 / 16.0, 7, 7, 0, L.AITB_F32S, 1, L.ptr(pooled), L.stream_ptr()))), 1)
res["fp32_us"] = round(t(lambda: ops.roi_align_forward(nhwc, rois.view(-1, 5), 1 / 16.0, 7, 7, 0, token_major=True)), 1)
nhwc_b = ops.transpose_cs(maps.reshape(B, 1024, -1), True, out_dtype=torch.bfloat16).view(B, 38, 63, 1024)
res["bf16_us"] = round(t(lambda: ops.roi_align_forward(nhwc_b, rois.view(-1, 5), 1 / 16.0, 7, 7, 0, token_major=True)), 1)
print(json.dumps(res))
