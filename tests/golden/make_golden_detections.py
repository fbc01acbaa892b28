"""Golden vectors for row f2 (detection post-processing) produced by EXECUTING the reference's own script lines.

    python tests/golden/make_golden_detections.py          (build container only: needs /root/reference)

The post-processing is not a function in the reference: it is the body of the evaluation loop of `test_net_voc.py`
(:380-450).  This script reads those source lines from /root/reference at run time (nothing is copied into the repo),
dedents them and `exec`s them UNMODIFIED, once per (image, query) unit, with the reference's own `cfg`
(model.utils.config), `bbox_transform_inv` / `clip_boxes` (model.rpn.bbox_transform) and `nms` (model.roi_layers.nms ->
the reference's C++ CPU kernel through oracle/_ref) in scope.  Only the environment is supplied: the loop variables the
block reads (`rois`, `cls_prob`, `bbox_pred`, `im_info`, `data`, `thresh`, `max_per_image`, `all_boxes`, `catgory`, `index`,
`args.class_agnostic`, `det_tic`) and `Tensor.cuda` as the identity (the block moves two constant tensors with `.cuda()`; there
is no GPU here).  The reference's CPU nms suppresses on IoU >= thr, the CUDA kernel it dispatches to on a GPU on IoU > thr
(SURVEY fact 3): the inputs below are generic floats with no exact IoU == 0.3 tie, so both give the same keep list.

Output: tests/golden/detections.pt -- inputs and, per unit, the reference's `all_boxes[catgory][index]`.
"""
import os
import sys
import textwrap
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from ait_b200 import synth  # noqa: E402
from oracle import ref_import  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
FIRST, LAST = 380, 450          # test_net_voc.py lines (1-based, inclusive)


def reference_block():
    path = os.path.join(ref_import.REF_ROOT, "test_net_voc.py")
    lines = open(path).read().split("\n")[FIRST - 1:LAST]
    assert lines[0].strip().startswith("scores = cls_prob.data"), lines[0]
    assert lines[-1].strip() == "pass", lines[-1]
    return compile(textwrap.dedent("\n".join(lines)), path + ":%d-%d" % (FIRST, LAST), "exec")


def run_case(code, rois, cls_prob, bbox_pred, im_info, thresh, max_per_image):
    from model.roi_layers import nms
    from model.rpn.bbox_transform import bbox_transform_inv, clip_boxes
    from model.utils.config import cfg
    out = []
    for b in range(rois.shape[0]):          # the reference evaluates one (image, query) per iteration, batch 1
        all_boxes = [[np.zeros((0, 5), dtype=np.float32)]]
        env = dict(torch=torch, np=np, time=time, cfg=cfg, nms=nms, bbox_transform_inv=bbox_transform_inv,
                   clip_boxes=clip_boxes, args=types.SimpleNamespace(class_agnostic=True),
                   rois=rois[b:b + 1].clone(), cls_prob=cls_prob[b:b + 1].clone(), bbox_pred=bbox_pred[b:b + 1].clone(),
                   im_info=im_info[b:b + 1].clone(), data=[None, None, im_info[b:b + 1].clone()],
                   thresh=thresh, max_per_image=max_per_image, all_boxes=all_boxes, catgory=0, index=0,
                   det_tic=time.time(), imdb=None)
        exec(code, env)
        out.append(torch.from_numpy(np.asarray(all_boxes[0][0], dtype=np.float32)).reshape(-1, 5).clone())
    return out


def main():
    ref_import.install()
    torch.Tensor.cuda = lambda self, *a, **k: self        # the block's two `.cuda()` calls on constant tensors
    code = reference_block()
    g = torch.Generator().manual_seed(31)
    B, P = 3, 300
    rois = torch.stack([synth.random_rois(u, P, batch_index=0) for u in range(B)])
    cls_prob = torch.rand(B, P, 1, generator=g)
    bbox_pred = 0.5 * torch.randn(B, P, 4, generator=g)
    im_info = torch.tensor([[600.0, 1000.0, 1.6], [600.0, 1000.0, 1.0], [600.0, 900.0, 0.75]])
    cases = []
    for thresh, max_per_image in ((0.0, 100), (0.5, 100), (0.0, 7), (0.999, 100), (0.0, 0)):
        dets = run_case(code, rois, cls_prob, bbox_pred, im_info, thresh, max_per_image)
        cases.append(dict(thresh=thresh, max_per_image=max_per_image, dets=dets))
        print("thresh %.3f max %3d -> detections per unit %s" % (thresh, max_per_image, [int(d.shape[0]) for d in dets]))
    torch.save(dict(rois=rois, cls_prob=cls_prob, bbox_pred=bbox_pred, im_info=im_info, nms_thresh=0.3, cases=cases,
                    source="test_net_voc.py:%d-%d" % (FIRST, LAST)), os.path.join(OUT, "detections.pt"))
    print("wrote", os.path.join(OUT, "detections.pt"), os.path.getsize(os.path.join(OUT, "detections.pt")))


if __name__ == "__main__":
    main()
