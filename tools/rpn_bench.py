"""Time the RPN head + proposal layer (rows f3a + f1 + a1/a2) on the benchmark shape: 8 units, C4 map 1024x38x63."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ait_b200 import synth  # noqa: E402
from ait_b200.rpn import _RPN  # noqa: E402

dev = "cuda:0"
B = 8
maps = torch.stack([synth.c4_map(u) for u in range(B)]).to(dev)
im_info = torch.tensor([[600.0, 1000.0, 1.0]] * B, device=dev)
for mode in (sys.argv[1:] or ["fp32", "tf32", "bf16"]):
    torch.manual_seed(0)
    m = _RPN(1024, compute_dtype=mode).to(dev).eval()
    with torch.no_grad():
        m.RPN_cls_score.weight.normal_(0, 0.05)
        m.RPN_bbox_pred.weight.normal_(0, 0.01)
    m._packed = None
    res = {}
    for name, fn in (("rpn_head", lambda: m.rpn_outputs(maps, im_info)), ("rpn_head+proposal_layer", lambda: m(maps, im_info))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st.record()
        for _ in range(10):
            fn()
        en.record()
        torch.cuda.synchronize()
        res[name + "_ms"] = round(st.elapsed_time(en) / 10, 4)
    flops = B * 38 * 63 * 2.0 * (512 * 9216 + 64 * 512)
    res.update(mode=mode, units=B, conv_tflops=round(flops / (res["rpn_head_ms"] * 1e-3) / 1e12, 1))
    print(json.dumps(res))
