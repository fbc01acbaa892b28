"""RPN proposal generation around the NMS (lib/model/rpn/proposal_layer.py:51-166,
bbox_transform.py:77-133, generate_anchors.py:45-105).

`propose_rois` is the hot-path tail: per-image descending top-N of the
scores, NMS (IoU > thr), first post_nms_topN survivors, zero-padded [B, post, 5] roi tensor --
three library calls for the whole batch, no python per-image loop, no host round trip.

Anchor enumeration / box decoding (the "next" row f1) are plain torch host-side helpers here; they
produce the synthetic RPN outputs for the benchmark and are not on the timed path.
"""
import numpy as np
import torch

from . import ops


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
