"""RPN proposal generation around the NMS (lib/model/rpn/proposal_layer.py:51-166,
bbox_transform.py:77-133, generate_anchors.py:45-105).

`propose_rois` is the hot-path tail: per-image descending top-N of the
scores, NMS (IoU > thr), first post_nms_topN survivors, zero-padded [B, post, 5] roi tensor --
three library calls for the whole batch, no python per-image loop, no host round trip.

`ProposalLayer` is the drop-in for the reference's `_ProposalLayer` (same constructor and input tuple): the
"next" row f1 -- anchor enumeration, bbox_transform_inv, clip_boxes and the NCHW re-ordering run in one
kernel (`aitb_rpn_decode`) in front of the top-n / NMS, so the whole layer is four library calls.

The torch helpers below (`shifted_anchors`, `decode_boxes`, `clip_boxes`) only produce the synthetic RPN
outputs of the benchmark; they are not on any timed path.
"""
import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from . import ops

# cfg[cfg_key].RPN_{PRE,POST}_NMS_TOP_N / RPN_NMS_THRESH (lib/model/utils/config.py:146-150,196-201)
RPN_CFG = {"TEST": dict(pre_nms_topN=6000, post_nms_topN=300, nms_thresh=0.7),
           "TRAIN": dict(pre_nms_topN=12000, post_nms_topN=2000, nms_thresh=0.7)}


def generate_anchors(base_size=16, ratios=(0.5, 1, 2), scales=(8, 16, 32)):
    """[A, 4] float64 anchors around the (0,0,base-1,base-1) window: for each ratio, for each scale."""
    ctr = 0.5 * (base_size - 1)
    area = float(base_size * base_size)
    out = []
    for r in ratios:
        w = np.round(np.sqrt(area / r))
        h = np.round(w * r)
        for s in scales:
            ws, hs = w * s, h * s
            out.append([ctr - 0.5 * (ws - 1), ctr - 0.5 * (hs - 1), ctr + 0.5 * (ws - 1), ctr + 0.5 * (hs - 1)])
    return np.asarray(out, dtype=np.float64)


def shifted_anchors(feat_h, feat_w, feat_stride=16, ratios=(0.5, 1, 2), scales=(8, 16, 32)):
    """[K*A, 4] float32, cells in row-major (y, x) order, A anchors per cell (proposal_layer.py:82-96)."""
    base = torch.from_numpy(generate_anchors(ratios=ratios, scales=scales)).float()
    sx = torch.arange(feat_w, dtype=torch.float32) * feat_stride
    sy = torch.arange(feat_h, dtype=torch.float32) * feat_stride
    yy, xx = torch.meshgrid(sy, sx, indexing="ij")
    shifts = torch.stack([xx.reshape(-1), yy.reshape(-1), xx.reshape(-1), yy.reshape(-1)], dim=1)
    return (base.view(1, -1, 4) + shifts.view(-1, 1, 4)).reshape(-1, 4)


def decode_boxes(anchors, deltas):
    """bbox_transform_inv (bbox_transform.py:77-106): anchors [N,4], deltas [B,N,4] -> [B,N,4]."""
    w = anchors[:, 2] - anchors[:, 0] + 1.0
    h = anchors[:, 3] - anchors[:, 1] + 1.0
    cx = anchors[:, 0] + 0.5 * w
    cy = anchors[:, 1] + 0.5 * h
    pcx = deltas[..., 0] * w + cx
    pcy = deltas[..., 1] * h + cy
    pw = torch.exp(deltas[..., 2]) * w
    ph = torch.exp(deltas[..., 3]) * h
    return torch.stack([pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw, pcy + 0.5 * ph], dim=-1)


def clip_boxes(boxes, im_h, im_w):
    """clip_boxes (bbox_transform.py:125-133) for a batch sharing one image size."""
    out = boxes.clone()
    out[..., 0::2] = out[..., 0::2].clamp(0, im_w - 1)
    out[..., 1::2] = out[..., 1::2].clamp(0, im_h - 1)
    return out


@L.on_tensor_device
def propose_rois(proposals, scores, pre_nms_topN=6000, post_nms_topN=300, nms_thresh=0.7):
    """proposals [B, K*A, 4], scores [B, K*A] (CUDA fp32) -> rois [B, post_nms_topN, 5], n_valid [B].

    Mirrors the loop of _ProposalLayer.forward (proposal_layer.py:129-166): sort descending, top
    pre_nms_topN, nms, first post_nms_topN, rows (i, x1, y1, x2, y2), zero padding."""
    if proposals.dim() != 3 or proposals.shape[2] != 4 or scores.shape != proposals.shape[:2]:
        raise RuntimeError("expected proposals [B, N, 4] and scores [B, N]")
    n_total = proposals.shape[1]
    n = n_total
    if 0 < pre_nms_topN < scores.numel():   # the reference's guard compares with the whole batch (:144)
        n = min(pre_nms_topN, n_total)
    order = ops.topk_desc(scores, n)
    _, n_keep, rois = ops.nms_batched(proposals, order, nms_thresh, post_nms_topN, mode=0, want_rois=True)
    return rois, n_keep


@L.on_tensor_device
def rpn_decode(rpn_cls_prob, rpn_bbox_pred, base_anchors, im_info, feat_stride):
    """rpn_cls_prob [B, 2A, H, W], rpn_bbox_pred [B, 4A, H, W], base_anchors [A, 4], im_info [B, 3]
    -> proposals [B, H*W*A, 4] (decoded, clipped), fg scores [B, H*W*A]  (proposal_layer.py:66-118)."""
    lib = L.load()
    ops._need_cuda(rpn_cls_prob, rpn_bbox_pred, base_anchors, im_info)
    scores = rpn_cls_prob.contiguous().float()
    deltas = rpn_bbox_pred.contiguous().float()
    B, c2, H, W = scores.shape
    A = c2 // 2
    if deltas.shape != (B, 4 * A, H, W) or base_anchors.shape != (A, 4) or im_info.shape != (B, 3):
        raise RuntimeError("rpn_decode: expected scores [B,2A,H,W], deltas [B,4A,H,W], anchors [A,4], im_info [B,3]")
    props = torch.empty((B, H * W * A, 4), dtype=torch.float32, device=scores.device)
    fg = torch.empty((B, H * W * A), dtype=torch.float32, device=scores.device)
    L.check(lib.aitb_rpn_decode(L.ptr(scores), L.ptr(deltas), L.ptr(base_anchors.contiguous().float()),
                                L.ptr(im_info.contiguous().float()), B, A, H, W, float(feat_stride), L.ptr(props),
                                L.ptr(fg), L.stream_ptr()))
    return props, fg


class ProposalLayer(nn.Module):
    """Drop-in for `_ProposalLayer` (lib/model/rpn/proposal_layer.py:28-166): forward(input) with
    input = (rpn_cls_prob [B,2A,H,W], rpn_bbox_pred [B,4A,H,W], im_info [B,3], cfg_key) -> rois [B, post_nms_topN, 5]."""

    def __init__(self, feat_stride, scales, ratios, cfg=None):
        super().__init__()
        self._feat_stride = feat_stride
        self.register_buffer("_anchors", torch.from_numpy(
            generate_anchors(scales=np.array(scales), ratios=np.array(ratios))).float(), persistent=False)
        self._num_anchors = self._anchors.size(0)
        self.cfg = cfg or RPN_CFG

    @L.on_tensor_device
    def forward(self, input):
        scores, bbox_deltas, im_info, cfg_key = input[0], input[1], input[2], input[3]
        c = self.cfg[cfg_key]
        proposals, fg = rpn_decode(scores, bbox_deltas, self._anchors.to(scores.device), im_info, self._feat_stride)
        rois, _ = propose_rois(proposals, fg, c["pre_nms_topN"], c["post_nms_topN"], c["nms_thresh"])
        return rois


# cfg.TRAIN.BBOX_NORMALIZE_{STDS,MEANS}, cfg.TEST.NMS (config.py:123-124,180)
BBOX_NORMALIZE_STDS = (0.1, 0.1, 0.2, 0.2)
BBOX_NORMALIZE_MEANS = (0.0, 0.0, 0.0, 0.0)


@L.on_tensor_device
def detections(rois, cls_prob, bbox_pred, im_info, thresh=0.0, nms_thresh=0.3, max_per_image=100,
               stds=BBOX_NORMALIZE_STDS, means=BBOX_NORMALIZE_MEANS, rescale=True):
    """Detection post-processing of test_net_voc.py:380-446 for a batch of units, on the device (row f2):
    de-normalise + decode + clip (+ rescale) the boxes, drop scores <= thresh, sort by score, NMS, keep the
    top max_per_image.  rois [B,P,5], cls_prob [B,P,1], bbox_pred [B,P,4], im_info [B,3]
    -> dets [B, P, 5] = (x1, y1, x2, y2, score) in descending score order (zero-padded), n_det [B] int32."""
    lib = L.load()
    ops._need_cuda(rois, cls_prob, bbox_pred, im_info)
    B, P = rois.shape[0], rois.shape[1]
    dev = rois.device
    rois = rois.contiguous().float()
    cls = cls_prob.reshape(B, P).contiguous().float()
    deltas = bbox_pred.reshape(B, P, 4).contiguous().float()
    pred = torch.empty((B, P, 4), dtype=torch.float32, device=dev)
    key = torch.empty((B, P), dtype=torch.float32, device=dev)
    n_valid = torch.empty((B,), dtype=torch.int32, device=dev)
    f4 = C.c_float * 4
    L.check(lib.aitb_box_decode(L.ptr(rois), 5, 1, L.ptr(deltas), L.ptr(cls), L.ptr(im_info.contiguous().float()), B, P,
                                f4(*stds), f4(*means), float(thresh), 1 if rescale else 0, L.ptr(pred), L.ptr(key),
                                L.ptr(n_valid), L.stream_ptr()))
    order = ops.topk_desc(key, P)
    keep, n_keep, _ = ops.nms_batched(pred, order, nms_thresh, P, mode=0)
    dets = torch.empty((B, P, 5), dtype=torch.float32, device=dev)
    n_det = torch.empty((B,), dtype=torch.int32, device=dev)
    L.check(lib.aitb_det_assemble(L.ptr(pred), L.ptr(cls), L.ptr(order), L.ptr(keep), L.ptr(n_keep), L.ptr(n_valid), B, P,
                                  int(max_per_image), L.ptr(dets), L.ptr(n_det), L.stream_ptr()))
    return dets, n_det
