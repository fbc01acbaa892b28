"""The hot path as one replayable unit: proposal-layer tail (top-n + NMS) -> ROIAlign -> AIT -> SKNet -> RCNN_top -> heads
(faster_rcnn_coatt_transformer_sk.py:247-337 between the RPN outputs and cls_prob / bbox_pred), optionally captured
into a CUDA graph.

A step is ~50 kernel launches on two streams (the proposal-independent query branch runs on the library's side stream).
At the benchmark batch (8 units x 300 proposals) the GPU is the bottleneck and launch cost hides behind it; at small
batches (1-2 units x 100 proposals: 1-1.5 ms of kernels) the ~50 launches plus the Python in between are a comparable
cost.  The library never allocates, never synchronises and has no device->host read on this path, so the whole step is
capturable as is: `DetectionPipeline(head, graph=True)` captures it once per input shape and replays it afterwards
(inputs are copied into the captured step's static buffers; the returned tensors are the step's static outputs and are
overwritten by the next call with the same shapes).
"""
import torch

from . import _lib as L
from .proposal import propose_rois


class DetectionPipeline:
    def __init__(self, head, pre_nms_topN=6000, post_nms_topN=300, nms_thresh=0.7, graph=True):
        self.head = head
        self.cfg = (int(pre_nms_topN), int(post_nms_topN), float(nms_thresh))
        self.graph = bool(graph)
        self._captured = {}

    def _step(self, non_img, non_qry, proposals, scores):
        rois, _ = propose_rois(proposals, scores, *self.cfg)
        cls_prob, bbox_pred = self.head.engine().head_forward(non_img, non_qry, rois)
        return rois, cls_prob, bbox_pred

    @L.on_tensor_device
    def __call__(self, non_img, non_qry, proposals, scores):
        """non_img [B,1024,H,W], non_qry [B,1024,8,8], proposals [B,N,4], scores [B,N] (CUDA fp32)
        -> rois [B,P,5], cls_prob [B,P,1], bbox_pred [B,P,4]."""
        if not self.graph:
            return self._step(non_img, non_qry, proposals, scores)
        eng = self.head.engine()
        key = (tuple(non_img.shape), tuple(proposals.shape), non_img.device, id(eng))
        ent = self._captured.get(key)
        if ent is None:
            static_in = [torch.empty_like(t) for t in (non_img, non_qry, proposals, scores)]
            for d, s in zip(static_in, (non_img, non_qry, proposals, scores)):
                d.copy_(s)
            side = torch.cuda.Stream(device=non_img.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):           # warm-up outside the capture: lazy attribute set-up, workspace growth
                for _ in range(2):
                    self._step(*static_in)
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                static_out = self._step(*static_in)
            ent = self._captured[key] = (g, static_in, static_out)
        g, static_in, static_out = ent
        for d, s in zip(static_in, (non_img, non_qry, proposals, scores)):
            if d.data_ptr() != s.data_ptr():
                d.copy_(s)
        g.replay()
        return static_out
