import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ait_b200 import ops
from oracle import c_ops

def rb(n, seed, span=400.0, size=150.0):
    g = torch.Generator().manual_seed(seed)
    xy = torch.rand(n, 2, generator=g) * span
    wh = torch.rand(n, 2, generator=g) * size + 1
    return torch.cat([xy, xy + wh], 1), torch.rand(n, generator=g)

for n, thr in [(129, 0.3), (129, 0.3), (129, 0.7), (200, 0.3), (3000, 0.3)]:
    boxes, scores = rb(n, 100 + n)
    order_ref = np.argsort(-scores.double().numpy(), kind="stable")
    order = ops.topk_desc(scores.cuda()[None], n)[0].cpu().numpy()
    print(n, thr, "topk ok", np.array_equal(order, order_ref))
    bs = boxes.numpy()[order_ref]
    ref_pos = c_ops.nms_sorted(bs, thr)
    keep, nk, _ = ops.nms_batched(torch.from_numpy(bs).cuda()[None], None, thr, n, mode=0)
    kp = keep[0, : int(nk)].cpu().numpy()
    print("  mode0 presorted ok", np.array_equal(kp, ref_pos), len(kp), len(ref_pos))
    if not np.array_equal(kp, ref_pos):
        a, b = set(kp.tolist()), set(ref_pos.tolist())
        print("  extra", sorted(a - b)[:20], "missing", sorted(b - a)[:20])
        for e in sorted(a - b)[:3]:
            for p in range(e):
                iou = None
                if p in b:
                    A, Bx = bs[p], bs[e]
                    l, r = max(A[0], Bx[0]), min(A[2], Bx[2]); t, bt = max(A[1], Bx[1]), min(A[3], Bx[3])
                    w, h = max(np.float32(r - l + 1), 0), max(np.float32(bt - t + 1), 0)
                    inter = np.float32(w * h)
                    sa = np.float32((A[2]-A[0]+1)*(A[3]-A[1]+1)); sb = np.float32((Bx[2]-Bx[0]+1)*(Bx[3]-Bx[1]+1))
                    iou = inter / (sa + sb - inter)
                    if iou > thr - 0.01:
                        print("   suppressor cand pos", p, "->", e, "iou", float(iou))
    keep1, nk1, _ = ops.nms_batched(boxes.cuda()[None], torch.from_numpy(order_ref).cuda()[None], thr, n, mode=1)
    k1 = keep1[0, : int(nk1)].cpu().numpy()
    print("  mode1 ok", np.array_equal(k1, c_ops.nms(boxes.numpy(), scores.numpy(), thr)))
