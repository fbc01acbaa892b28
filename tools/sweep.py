"""BASELINE config 5: throughput sweep over proposals per unit (100..1000) x queries per image (1..16).

    python tools/sweep.py [fp32|tf32|bf16] [images]

Every (image, query) is one unit with its own C4 map and RPN outputs (co-attention runs before the RPN in the
reference, faster_rcnn_coatt_transformer_sk.py:234-247).  One line per point: ms per batch and pairs / s on this
GPU (device-resident inputs, CUDA events, 3 warm-up + 5 timed passes).
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ait_b200 import synth  # noqa: E402
from ait_b200.proposal import propose_rois  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
images = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
head = synth.make_head(seed=0, calibrated=True, randomize_bn=True, compute_dtype=mode).to(dev)
eng = head.engine()
max_units = images * 16
maps = torch.stack([synth.c4_map(u) for u in range(max_units)]).to(dev)
qrys = torch.stack([synth.query_feat(u) for u in range(max_units)]).to(dev)
rpn = [synth.rpn_outputs(u) for u in range(max_units)]
boxes = torch.stack([r[0] for r in rpn]).to(dev)
scores = torch.stack([r[1] for r in rpn]).to(dev)
rows = []
for P in (100, 200, 300, 500, 1000):
    for Q in (1, 2, 4, 8, 16):
        U = images * Q

        def step():
            rois, _ = propose_rois(boxes[:U], scores[:U], 6000, P, 0.7)
            return eng.head_forward(maps[:U], qrys[:U], rois)

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st.record()
        for _ in range(5):
            cls, _ = step()
        en.record()
        torch.cuda.synchronize()
        ms = st.elapsed_time(en) / 5
        row = dict(mode=mode, images=images, queries=Q, proposals=P, units=U, pairs=U * P, ms=round(ms, 3),
                   pairs_per_s=round(U * P / (ms * 1e-3)), finite=bool(torch.isfinite(cls).all()))
        rows.append(row)
        print(json.dumps(row), flush=True)
