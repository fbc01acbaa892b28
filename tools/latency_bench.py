"""Small-batch latency of the hot path, eager launches vs one CUDA-graph replay (ait_b200.pipeline.DetectionPipeline):
    python tools/latency_bench.py [fp32|bf16]
ms per step (proposal top-n + NMS + head), device-resident inputs, CUDA events, 5 warm-up + 30 timed steps."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ait_b200 import synth  # noqa: E402
from ait_b200.pipeline import DetectionPipeline  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
dev = "cuda:0"
head = synth.spread_score_layer(synth.make_head(seed=0, calibrated=True, randomize_bn=True, compute_dtype=mode)).to(dev)


def t(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return round(s.elapsed_time(e) / n, 4)


for U, P in ((1, 100), (2, 100), (1, 300), (2, 300), (8, 300)):
    maps = torch.stack([synth.c4_map(u) for u in range(U)]).to(dev)
    qrys = torch.stack([synth.query_feat(u) for u in range(U)]).to(dev)
    rpn = [synth.rpn_outputs(u) for u in range(U)]
    boxes, scores = torch.stack([r[0] for r in rpn]).to(dev), torch.stack([r[1] for r in rpn]).to(dev)
    eager = DetectionPipeline(head, 6000, P, 0.7, graph=False)
    graph = DetectionPipeline(head, 6000, P, 0.7, graph=True)
    ms_e = t(lambda: eager(maps, qrys, boxes, scores))
    ms_g = t(lambda: graph(maps, qrys, boxes, scores))
    print(json.dumps(dict(mode=mode, units=U, proposals=P, pairs=U * P, eager_ms=ms_e, graph_ms=ms_g,
                          eager_pairs_per_s=round(U * P / (ms_e * 1e-3)), graph_pairs_per_s=round(U * P / (ms_g * 1e-3)))))
