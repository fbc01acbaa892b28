"""Eager launches vs one CUDA-graph replay of the whole step (proposal top-n + NMS + head), benchmark shape."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ait_b200 import synth  # noqa: E402
from ait_b200.proposal import propose_rois  # noqa: E402

dev = torch.device("cuda:0")
B, P = 8, 300
for mode in (sys.argv[1:] or ["fp32", "bf16"]):
    head = synth.make_head(seed=0, calibrated=True, randomize_bn=True, compute_dtype=mode).to(dev)
    eng = head.engine()
    maps = torch.stack([synth.c4_map(u) for u in range(B)]).to(dev)
    qrys = torch.stack([synth.query_feat(u) for u in range(B)]).to(dev)
    rpn = [synth.rpn_outputs(u) for u in range(B)]
    boxes = torch.stack([r[0] for r in rpn]).to(dev)
    scores = torch.stack([r[1] for r in rpn]).to(dev)

    def step():
        rois, _ = propose_rois(boxes, scores, 6000, P, 0.7)
        return eng.head_forward(maps, qrys, rois)

    def timeit(fn, n=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st.record()
        for _ in range(n):
            fn()
        en.record()
        torch.cuda.synchronize()
        return st.elapsed_time(en) / n

    eager = timeit(step)
    ref = step()[0].clone()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        step()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = step()
    graphed = timeit(g.replay)
    g.replay()
    torch.cuda.synchronize()
    print(json.dumps({"mode": mode, "eager_ms": round(eager, 3), "graph_ms": round(graphed, 3),
                      "same": bool(torch.equal(out[0], ref))}))
