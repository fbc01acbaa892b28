"""Training-only samplers and losses around the head, on the device.

Drop-ins for the reference's
  `_AnchorTargetLayer`   (lib/model/rpn/anchor_target_layer.py:33-199)          -> `AnchorTargetLayer`
  `_ProposalTargetLayer` (lib/model/rpn/proposal_target_layer_cascade.py:20-220) -> `ProposalTargetLayer`
  the RPN loss lines of `_RPN.forward` (lib/model/rpn/rpn.py:99-126)            -> `rpn_losses`
  the detection-loss lines of `_fasterRCNN.forward`
  (lib/model/faster_rcnn/faster_rcnn_coatt_transformer_sk.py:334-361)           -> `rcnn_losses`
  `_smooth_l1_loss` (lib/model/utils/net_utils.py:75-89) lives inside both loss kernels.

Random sub-sampling.  The reference draws from numpy's GLOBAL generator with data-dependent lengths
(`np.random.permutation(fg_inds.size(0))`, ...), which forces one device->host read of the per-image candidate
counts (the reference pays the same synchronisation in `.numel()` / `nonzero`).  `rng="numpy"` (default) makes
exactly the reference's calls in the reference's order, so after `np.random.seed(s)` the layers pick the same
anchors and rois as the reference, bit for bit.  The candidate lists never leave the device: the host only sends
RANKS (positions in the ascending candidate lists), 128 ints or a byte mask per image.
`rng=<np.random.Generator>` draws from a private generator instead (same distribution, independent stream).
`rng="device"` keeps the whole layer on the device: the selections are drawn by a counter-based generator inside the
library (Philox4x32-10 keyed by `seed`, the call counter, the image and the candidate), so there is NO device->host
read and no synchronisation -- the mode for training throughput when bit parity with numpy's stream is not needed.

There is no CPU fallback: CPU tensors raise.
"""
import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from .proposal import generate_anchors

# cfg.TRAIN.* as the reference trains: lib/model/utils/config.py:23,81-158 merged with cfgs/res50.yml (trainval_net_voc.py:206-209
# always loads one of cfgs/res50*.yml / res101.yml; all of them override BG_THRESH_LO to 0.0 -- config.py's 0.1 is never in effect)
TRAIN_CFG = dict(BATCH_SIZE=128, FG_FRACTION=0.25, FG_THRESH=0.5, BG_THRESH_HI=0.5, BG_THRESH_LO=0.0,
                 BBOX_NORMALIZE_MEANS=(0.0, 0.0, 0.0, 0.0), BBOX_NORMALIZE_STDS=(0.1, 0.1, 0.2, 0.2),
                 BBOX_INSIDE_WEIGHTS=(1.0, 1.0, 1.0, 1.0), BBOX_NORMALIZE_TARGETS_PRECOMPUTED=True,
                 RPN_POSITIVE_OVERLAP=0.7, RPN_NEGATIVE_OVERLAP=0.3, RPN_CLOBBER_POSITIVES=False,
                 RPN_FG_FRACTION=0.5, RPN_BATCHSIZE=256, RPN_BBOX_INSIDE_WEIGHTS=(1.0, 1.0, 1.0, 1.0),
                 RPN_POSITIVE_WEIGHT=-1.0, MARGIN=-0.3)


def _rng(rng):
    if rng == "numpy":
        return np.random           # the module-level functions share the global MT19937 stream with the reference
    if isinstance(rng, np.random.Generator):
        class _G:                  # same call names on a private generator
            permutation = staticmethod(rng.permutation)
            rand = staticmethod(lambda n: rng.random(n))
        return _G
    raise RuntimeError("rng must be \"numpy\" (the reference's global stream), a numpy.random.Generator or \"device\"")


class AnchorTargetLayer(nn.Module):
    """forward(input) with input = (rpn_cls_score [B,2A,H,W], gt_boxes [B,K,5], im_info [B,3], num_boxes) ->
    [labels [B,1,A*H,W], bbox_targets [B,4A,H,W], bbox_inside_weights, bbox_outside_weights] (CUDA fp32)."""

    def __init__(self, feat_stride, scales, ratios, cfg=None, rng="numpy", seed=0):
        super().__init__()
        self.seed, self._calls = int(seed), 0
        self._feat_stride = feat_stride
        self._scales = scales
        self.register_buffer("_anchors", torch.from_numpy(
            generate_anchors(scales=np.array(scales), ratios=np.array(ratios))).float(), persistent=False)
        self._num_anchors = self._anchors.size(0)
        self._allowed_border = 0
        self.cfg = dict(TRAIN_CFG, **(cfg or {}))
        self.rng = rng
        if self.cfg["RPN_POSITIVE_WEIGHT"] >= 0:
            raise RuntimeError("AnchorTargetLayer: only the uniform weighting (RPN_POSITIVE_WEIGHT < 0) exists in the "
                               "reference (anchor_target_layer.py:162-171 leaves the other branch undefined)")

    @L.on_tensor_device
    def forward(self, input):
        rpn_cls_score, gt_boxes, im_info = input[0], input[1], input[2]
        lib = L.load()
        ops._need_cuda(gt_boxes, im_info)
        cfg = self.cfg
        dev = gt_boxes.device
        H, W = int(rpn_cls_score.size(2)), int(rpn_cls_score.size(3))
        A = self._num_anchors
        gt = gt_boxes.contiguous().float()
        info = im_info.contiguous().float()
        B, K = gt.shape[0], gt.shape[1]
        total = A * H * W
        base = self._anchors.to(dev)
        labels = torch.empty((B, total), dtype=torch.int8, device=dev)
        argmax = torch.empty((B, total), dtype=torch.int32, device=dev)
        counts = torch.empty((B, 2), dtype=torch.int32, device=dev)
        ws_bytes = lib.aitb_anchor_target_workspace_bytes(B, A, H, W, K)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        L.check(lib.aitb_anchor_target_assign(L.ptr(base), L.ptr(gt), L.ptr(info), B, A, H, W, K, float(self._feat_stride),
                                              float(cfg["RPN_NEGATIVE_OVERLAP"]), float(cfg["RPN_POSITIVE_OVERLAP"]),
                                              1 if cfg["RPN_CLOBBER_POSITIVES"] else 0, L.ptr(labels), L.ptr(argmax),
                                              L.ptr(counts), L.ptr(ws), ws_bytes, L.stream_ptr()))
        num_fg = int(cfg["RPN_FG_FRACTION"] * cfg["RPN_BATCHSIZE"])
        if self.rng == "device":
            self._calls += 1
            L.check(lib.aitb_anchor_target_subsample_device(L.ptr(labels), L.ptr(counts), B, total, num_fg, int(cfg["RPN_BATCHSIZE"]),
                                                            C.c_uint64((self.seed * 0x9E3779B97F4A7C15 + self._calls) & (2 ** 64 - 1)),
                                                            L.stream_ptr()))
            return self._finish(lib, base, gt, B, A, H, W, K, labels, argmax, None, 1, cfg, dev)
        # the one host round trip: (#fg, #bg) per image decide how many random numbers the reference draws (:130-152)
        cnt = counts.cpu().numpy()
        rnd = _rng(self.rng)
        ld = int(max(1, cnt.max()))
        drop = np.zeros((B, 2, ld), dtype=np.uint8)
        any_drop = False
        for i in range(B):
            n_f, n_b = int(cnt[i, 0]), int(cnt[i, 1])
            fg_left = n_f
            if n_f > num_fg:
                perm = rnd.permutation(n_f)
                drop[i, 0, perm[:n_f - num_fg]] = 1
                fg_left = num_fg
                any_drop = True
            num_bg = cfg["RPN_BATCHSIZE"] - fg_left
            if n_b > num_bg:
                perm = rnd.permutation(n_b)
                drop[i, 1, perm[:n_b - num_bg]] = 1
                any_drop = True
        drop_d = torch.from_numpy(drop).to(dev, non_blocking=False) if any_drop else None
        return self._finish(lib, base, gt, B, A, H, W, K, labels, argmax, drop_d, ld, cfg, dev)

    def _finish(self, lib, base, gt, B, A, H, W, K, labels, argmax, drop_d, ld, cfg, dev):
        n_examples = torch.empty((B,), dtype=torch.int32, device=dev)
        labels_out = torch.empty((B, 1, A * H, W), dtype=torch.float32, device=dev)
        targets = torch.empty((B, 4 * A, H, W), dtype=torch.float32, device=dev)
        inside = torch.empty_like(targets)
        outside = torch.empty_like(targets)
        L.check(lib.aitb_anchor_target_finish(L.ptr(base), L.ptr(gt), B, A, H, W, K, float(self._feat_stride), L.ptr(labels),
                                              L.ptr(argmax), L.ptr(drop_d), ld, float(cfg["RPN_BBOX_INSIDE_WEIGHTS"][0]),
                                              L.ptr(n_examples), L.ptr(labels_out), L.ptr(targets), L.ptr(inside),
                                              L.ptr(outside), L.stream_ptr()))
        return [labels_out, targets, inside, outside]


class ProposalTargetLayer(nn.Module):
    """forward(all_rois [B,R,5], gt_boxes [B,K,5], num_boxes) -> rois [B,128,5], labels [B,128], bbox_targets,
    bbox_inside_weights, bbox_outside_weights [B,128,4] (CUDA fp32)."""

    def __init__(self, nclasses, cfg=None, rng="numpy", seed=0):
        super().__init__()
        self.seed, self._calls = int(seed), 0
        self._num_classes = nclasses
        self.cfg = dict(TRAIN_CFG, **(cfg or {}))
        self.rng = rng

    @L.on_tensor_device
    def forward(self, all_rois, gt_boxes, num_boxes=None):
        lib = L.load()
        ops._need_cuda(all_rois, gt_boxes)
        cfg = self.cfg
        dev = gt_boxes.device
        rois = all_rois.contiguous().float()
        gt = gt_boxes.contiguous().float()
        B, R, K = rois.shape[0], rois.shape[1], gt.shape[1]
        if rois.shape[2] != 5 or gt.shape[2] != 5 or gt.shape[0] != B:
            raise RuntimeError("ProposalTargetLayer: expected all_rois [B,R,5] and gt_boxes [B,K,5]")
        N = R + K
        S = int(cfg["BATCH_SIZE"] / 1)                                   # num_images = 1 (:47-48)
        fg_per_image = int(np.round(cfg["FG_FRACTION"] * S)) or 1
        max_ov = torch.empty((B, N), dtype=torch.float32, device=dev)
        assign = torch.empty((B, N), dtype=torch.int32, device=dev)
        cls = torch.empty((B, N), dtype=torch.int8, device=dev)
        counts = torch.empty((B, 2), dtype=torch.int32, device=dev)
        L.check(lib.aitb_proposal_target_assign(L.ptr(rois), L.ptr(gt), B, R, K, float(cfg["FG_THRESH"]),
                                                float(cfg["BG_THRESH_HI"]), float(cfg["BG_THRESH_LO"]), L.ptr(max_ov),
                                                L.ptr(assign), L.ptr(cls), L.ptr(counts), L.stream_ptr()))
        if self.rng == "device":
            self._calls += 1
            picks_d = torch.empty((B, S), dtype=torch.int32, device=dev)
            nfp_d = torch.empty((B,), dtype=torch.int32, device=dev)
            bad = torch.zeros((1,), dtype=torch.int32, device=dev)
            L.check(lib.aitb_proposal_target_picks_device(L.ptr(counts), B, S, fg_per_image,
                                                          C.c_uint64((self.seed * 0x9E3779B97F4A7C15 + self._calls) & (2 ** 64 - 1)),
                                                          L.ptr(picks_d), L.ptr(nfp_d), L.ptr(bad), L.stream_ptr()))
            return self._sample(lib, rois, gt, B, R, K, N, S, cls, assign, picks_d, nfp_d, bad, cfg, dev)
        cnt = counts.cpu().numpy()                                        # the one host round trip (see module doc)
        rnd = _rng(self.rng)
        picks = np.zeros((B, S), dtype=np.int32)
        n_fg_pick = np.zeros((B,), dtype=np.int32)
        for i in range(B):                                                # :154-199, the same draws in the same order
            nf, nb = int(cnt[i, 0]), int(cnt[i, 1])
            if nf > 0 and nb > 0:
                fg_this = min(fg_per_image, nf)
                picks[i, :fg_this] = rnd.permutation(nf)[:fg_this]
                picks[i, fg_this:] = np.floor(rnd.rand(S - fg_this) * nb)
            elif nf > 0:
                fg_this = S
                picks[i] = np.floor(rnd.rand(S) * nf)
            elif nb > 0:
                fg_this = 0
                picks[i] = np.floor(rnd.rand(S) * nb)
            else:
                raise ValueError("bg_num_rois = 0 and fg_num_rois = 0, this should not happen!")
            n_fg_pick[i] = fg_this
        picks_d = torch.from_numpy(picks).to(dev)
        nfp_d = torch.from_numpy(n_fg_pick).to(dev)
        bad = torch.zeros((1,), dtype=torch.int32, device=dev)
        return self._sample(lib, rois, gt, B, R, K, N, S, cls, assign, picks_d, nfp_d, bad, cfg, dev)

    def _sample(self, lib, rois, gt, B, R, K, N, S, cls, assign, picks_d, nfp_d, bad, cfg, dev):
        lists = torch.empty((B, 2, N), dtype=torch.int32, device=dev)
        rois_out = torch.zeros((B, S, 5), dtype=torch.float32, device=dev)
        labels = torch.zeros((B, S), dtype=torch.float32, device=dev)
        targets = torch.zeros((B, S, 4), dtype=torch.float32, device=dev)
        inside = torch.zeros_like(targets)
        outside = torch.zeros_like(targets)
        f4 = C.c_float * 4
        means = cfg["BBOX_NORMALIZE_MEANS"] if cfg["BBOX_NORMALIZE_TARGETS_PRECOMPUTED"] else (0.0,) * 4
        stds = cfg["BBOX_NORMALIZE_STDS"] if cfg["BBOX_NORMALIZE_TARGETS_PRECOMPUTED"] else (1.0,) * 4
        L.check(lib.aitb_proposal_target_sample(L.ptr(rois), L.ptr(gt), B, R, K, L.ptr(cls), L.ptr(assign), L.ptr(picks_d),
                                                L.ptr(nfp_d), S, f4(*means), f4(*stds), f4(*cfg["BBOX_INSIDE_WEIGHTS"]),
                                                L.ptr(lists), L.ptr(rois_out), L.ptr(labels), L.ptr(targets), L.ptr(inside),
                                                L.ptr(outside), L.ptr(bad), L.stream_ptr()))
        # device flag: non-zero if a pick fell outside its list (never with our draws) or, in the "device" mode, if an
        # image had neither foreground nor background candidates (where the host modes raise like the reference)
        self.last_bad_flag = bad
        return rois_out, labels, targets, inside, outside


_AnchorTargetLayer = AnchorTargetLayer
_ProposalTargetLayer = ProposalTargetLayer


class _RPNLossFn(torch.autograd.Function):
    @staticmethod
    @L.on_tensor_device
    def forward(ctx, rpn_cls_score, rpn_bbox_pred, labels, targets, inside, outside, sigma):
        lib = L.load()
        ops._need_cuda(rpn_cls_score, rpn_bbox_pred, labels, targets, inside, outside)
        B, c2, H, W = rpn_cls_score.shape
        A = c2 // 2
        if rpn_bbox_pred.shape != (B, 4 * A, H, W) or labels.numel() != B * A * H * W or targets.shape != rpn_bbox_pred.shape:
            raise RuntimeError("rpn_losses: expected rpn_cls_score [B,2A,H,W], rpn_bbox_pred [B,4A,H,W], labels [B,1,A*H,W]")
        t = [x.detach().contiguous().float() for x in (rpn_cls_score, rpn_bbox_pred, labels, targets, inside, outside)]
        losses = torch.empty((2,), dtype=torch.float32, device=t[0].device)
        acc = torch.empty((3,), dtype=torch.float64, device=t[0].device)
        L.check(lib.aitb_rpn_loss(*[L.ptr(x) for x in t], B, A, H, W, float(sigma), None, L.ptr(losses), None, None,
                                  L.ptr(acc), L.stream_ptr()))
        ctx.save_for_backward(*t)
        ctx.meta = (B, A, H, W, float(sigma))
        return losses[0], losses[1]

    @staticmethod
    @L.on_tensor_device
    def backward(ctx, g_cls, g_box):
        lib = L.load()
        t = ctx.saved_tensors
        B, A, H, W, sigma = ctx.meta
        dev = t[0].device
        gscale = torch.stack([g_cls.reshape(()).float(), g_box.reshape(()).float()]).contiguous()
        losses = torch.empty((2,), dtype=torch.float32, device=dev)
        acc = torch.empty((3,), dtype=torch.float64, device=dev)
        d_score, d_bbox = torch.empty_like(t[0]), torch.empty_like(t[1])
        L.check(lib.aitb_rpn_loss(*[L.ptr(x) for x in t], B, A, H, W, sigma, L.ptr(gscale), L.ptr(losses), L.ptr(d_score),
                                  L.ptr(d_bbox), L.ptr(acc), L.stream_ptr()))
        return d_score, d_bbox, None, None, None, None, None


def rpn_losses(rpn_cls_score, rpn_bbox_pred, rpn_data, sigma=3.0):
    """rpn.py:99-126: (rpn_loss_cls, rpn_loss_box) from the raw RPN outputs (rpn_cls_score [B,2A,H,W] with the
    background channels first, rpn_bbox_pred [B,4A,H,W]) and the anchor-target outputs `rpn_data`."""
    labels, targets, inside, outside = rpn_data
    return _RPNLossFn.apply(rpn_cls_score, rpn_bbox_pred, labels, targets, inside, outside, sigma)


class _RCNNLossFn(torch.autograd.Function):
    @staticmethod
    @L.on_tensor_device
    def forward(ctx, score, bbox_pred, labels, targets, inside, outside, bs, margin, margin_scale):
        lib = L.load()
        ops._need_cuda(score, bbox_pred, labels, targets, inside, outside)
        n = score.shape[0]
        if score.shape != (n, 2) or bbox_pred.shape != (n, 4) or labels.numel() != n or n % bs:
            raise RuntimeError("rcnn_losses: expected score [bs*P,2], bbox_pred [bs*P,4], rois_label [bs*P]")
        P = n // bs
        t = [score.detach().contiguous().float(), bbox_pred.detach().contiguous().float(),
             labels.detach().reshape(-1).contiguous().float(), targets.detach().reshape(n, 4).contiguous().float(),
             inside.detach().reshape(n, 4).contiguous().float(), outside.detach().reshape(n, 4).contiguous().float()]
        dev = t[0].device
        losses = torch.empty((3,), dtype=torch.float32, device=dev)
        acc = torch.empty((3,), dtype=torch.float64, device=dev)
        L.check(lib.aitb_rcnn_loss(*[L.ptr(x) for x in t], bs, P, float(margin), float(margin_scale), None, L.ptr(losses),
                                   None, None, None, L.ptr(acc), L.stream_ptr()))
        ctx.save_for_backward(*t)
        ctx.meta = (bs, P, float(margin), float(margin_scale))
        return losses[0], losses[1], losses[2]

    @staticmethod
    @L.on_tensor_device
    def backward(ctx, g_cls, g_margin, g_bbox):
        lib = L.load()
        t = ctx.saved_tensors
        bs, P, margin, margin_scale = ctx.meta
        dev = t[0].device
        gscale = torch.stack([g_cls.reshape(()).float(), g_margin.reshape(()).float(), g_bbox.reshape(()).float()]).contiguous()
        losses = torch.empty((3,), dtype=torch.float32, device=dev)
        acc = torch.empty((3,), dtype=torch.float64, device=dev)
        d_score, d_bbox = torch.empty_like(t[0]), torch.empty_like(t[1])
        L.check(lib.aitb_rcnn_loss(*[L.ptr(x) for x in t], bs, P, margin, margin_scale, L.ptr(gscale), L.ptr(losses), None,
                                   L.ptr(d_score), L.ptr(d_bbox), L.ptr(acc), L.stream_ptr()))
        return d_score, d_bbox, None, None, None, None, None, None, None


def rcnn_losses(score, bbox_pred, rois_label, rois_target, rois_inside_ws, rois_outside_ws, bs, margin=TRAIN_CFG["MARGIN"],
                margin_scale=3.0):
    """faster_rcnn_coatt_transformer_sk.py:340-361: (RCNN_loss_cls, margin_loss, RCNN_loss_bbox) from score [bs*P,2],
    bbox_pred [bs*P,4] and the proposal-target outputs."""
    return _RCNNLossFn.apply(score, bbox_pred, rois_label, rois_target, rois_inside_ws, rois_outside_ws, int(bs),
                             margin, margin_scale)


class _ScoreHeadsFn(torch.autograd.Function):
    """score / bbox heads on the pooled features, differentiable w.r.t. both feature inputs and the six parameters."""

    @staticmethod
    @L.on_tensor_device
    def forward(ctx, feat, qfeat, P, w_bbox, b_bbox, w1, b1, w2, b2):
        lib = L.load()
        ops._need_cuda(feat, qfeat, w_bbox, b_bbox, w1, b1, w2, b2)
        G = feat.shape[0]
        if feat.shape != (G, 2048) or G % P or qfeat.shape != (G // P, 2048) or w1.shape != (8, 4096) or w2.shape != (2, 8) \
                or w_bbox.shape != (4, 2048):
            raise RuntimeError("score_heads: expected feat [G,2048], qfeat [G/P,2048], RCNN_bbox_pred Linear(2048,4), "
                               "RCNN_cls_score = Linear(4096,8), Linear(8,2)")
        t = [x.detach().contiguous().float() for x in (feat, qfeat, w_bbox, b_bbox, w1, b1, w2, b2)]
        dev = t[0].device
        bbox = torch.empty((G, 4), dtype=torch.float32, device=dev)
        hidden = torch.empty((G, 8), dtype=torch.float32, device=dev)
        score = torch.empty((G, 2), dtype=torch.float32, device=dev)
        L.check(lib.aitb_heads_forward_train(L.ptr(t[0]), L.ptr(t[1]), G, P, *[L.ptr(x) for x in t[2:]], L.ptr(bbox),
                                             L.ptr(hidden), L.ptr(score), L.stream_ptr()))
        ctx.save_for_backward(t[0], t[1], hidden, t[2], t[4], t[6])
        ctx.P = P
        return score, bbox

    @staticmethod
    @L.on_tensor_device
    def backward(ctx, d_score, d_bbox):
        lib = L.load()
        feat, qfeat, hidden, w_bbox, w1, w2 = ctx.saved_tensors
        G, P, dev = feat.shape[0], ctx.P, feat.device
        d_score = (torch.zeros((G, 2), device=dev) if d_score is None else d_score).contiguous().float()
        d_bbox = (torch.zeros((G, 4), device=dev) if d_bbox is None else d_bbox).contiguous().float()
        d_feat, d_qfeat = torch.empty_like(feat), torch.empty_like(qfeat)
        g = [torch.zeros(s, dtype=torch.float32, device=dev) for s in ((4, 2048), (4,), (8, 4096), (8,), (2, 8), (2,))]
        nb = lib.aitb_heads_backward_workspace_bytes(G, P)
        ws = torch.empty((nb,), dtype=torch.uint8, device=dev)
        L.check(lib.aitb_heads_backward(L.ptr(feat), L.ptr(qfeat), L.ptr(hidden), L.ptr(d_score), L.ptr(d_bbox), G, P, L.ptr(w_bbox),
                                        L.ptr(w1), L.ptr(w2), L.ptr(d_feat), L.ptr(d_qfeat), *[L.ptr(x) for x in g], L.ptr(ws), nb,
                                        L.stream_ptr()))
        return d_feat, d_qfeat, None, g[0], g[1], g[2], g[3], g[4], g[5]


def score_heads(feat, qfeat, P, RCNN_bbox_pred, RCNN_cls_score):
    """faster_rcnn_coatt_transformer_sk.py:318-332 in training: (score [G,2] logits, bbox_pred [G,4]) from the pooled
    layer-4 features feat [G,2048] of the pairs and qfeat [G/P,2048] of the units' queries; RCNN_bbox_pred = nn.Linear(2048, 4),
    RCNN_cls_score = nn.Sequential(nn.Linear(4096, 8), nn.Linear(8, 2)) (the reference's modules, ordinary Parameters).
    Differentiable: feeds `rcnn_losses`, back-propagates to both feature inputs and the six parameters."""
    return _ScoreHeadsFn.apply(feat, qfeat, int(P), RCNN_bbox_pred.weight, RCNN_bbox_pred.bias, RCNN_cls_score[0].weight,
                               RCNN_cls_score[0].bias, RCNN_cls_score[1].weight, RCNN_cls_score[1].bias)


@L.on_tensor_device
def mean_pool_backward(d_feat):
    """Adjoint of `_head_to_tail`'s spatial mean: d_feat [G,2048] -> d_top [G,16,2048] (token-major 4x4 map)."""
    lib = L.load()
    ops._need_cuda(d_feat)
    d_feat = d_feat.contiguous().float()
    G = d_feat.shape[0]
    d_top = torch.empty((G, 16, 2048), dtype=torch.float32, device=d_feat.device)
    L.check(lib.aitb_mean_pool_backward(L.ptr(d_feat), G, L.ptr(d_top), L.stream_ptr()))
    return d_top
