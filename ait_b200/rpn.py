"""Drop-in for the reference's `_RPN` (lib/model/rpn/rpn.py:18-110), inference path: RPN_Conv 3x3 + ReLU,
RPN_cls_score | RPN_bbox_pred, the bg/fg softmax and the whole proposal layer, executed by libaitb200
(`aitb_rpn_forward`: one tcgen05 conv GEMM over the C4 map with nine shifted TMA boxes per K chunk, one fused
1x1 GEMM for both heads, one decode kernel) followed by the on-device top-n + NMS.  Same parameter names as
the reference (RPN_Conv.*, RPN_cls_score.*, RPN_bbox_pred.*), so detector checkpoints load unchanged.
`.train()` runs the training branch of rpn.py:85-140 on the device: the differentiable head of ait_b200/rpn_train.py (own
backward), the proposal layer with the TRAIN settings, `AnchorTargetLayer` and the two RPN losses (ait_b200/targets.py).
"""
import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from .packing import fingerprint, round_to_tf32
from .proposal import RPN_CFG, generate_anchors, propose_rois


class RPNWeights(C.Structure):
    _fields_ = [("dtype", C.c_int), ("round_tf32", C.c_int), ("A", C.c_int), ("n_pad", C.c_int),
                ("conv", L.Linear), ("heads", L.Linear)]


class _RPN(nn.Module):
    def __init__(self, din, anchor_scales=(8, 16, 32), anchor_ratios=(0.5, 1, 2), feat_stride=16, cfg=None,
                 compute_dtype=torch.float32):
        super().__init__()
        if din != 1024:
            raise NotImplementedError("ait_b200._RPN: ResNet-50 C4 input (1024 channels) only")
        self.din = din
        self.anchor_scales, self.anchor_ratios, self.feat_stride = anchor_scales, anchor_ratios, feat_stride
        A = len(anchor_scales) * len(anchor_ratios)
        self.nc_score_out, self.nc_bbox_out = 2 * A, 4 * A
        self.RPN_Conv = nn.Conv2d(din, 512, 3, 1, 1, bias=True)
        self.RPN_cls_score = nn.Conv2d(512, self.nc_score_out, 1, 1, 0)
        self.RPN_bbox_pred = nn.Conv2d(512, self.nc_bbox_out, 1, 1, 0)
        self.register_buffer("_anchors", torch.from_numpy(
            generate_anchors(scales=np.array(anchor_scales), ratios=np.array(anchor_ratios))).float(), persistent=False)
        self.cfg = cfg or RPN_CFG
        self.compute_dtype = compute_dtype
        self._packed = None
        self.rpn_loss_cls = 0
        self.rpn_loss_box = 0
        self._anchor_target = None      # built on first use in .train() (ait_b200.targets.AnchorTargetLayer, not a sub-module:
        #                                 the reference's RPN_anchor_target has no parameters and no state_dict entries either)

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def _load_from_state_dict(self, *a, **k):
        super()._load_from_state_dict(*a, **k)
        self._packed = None

    def _pack(self):
        mode = L.mode_name(self.compute_dtype)
        A = self.nc_score_out // 2
        n_pad = 64 if 6 * A <= 64 else 128
        if 6 * A > 128:
            raise NotImplementedError("ait_b200._RPN: at most 21 anchors per cell")

        def mat(t):
            t = t.detach().float().contiguous()
            if mode == "tf32":
                return round_to_tf32(t)
            if mode == "bf16":
                return t.to(torch.bfloat16)
            return ops.split_planes(t)

        with torch.no_grad():
            wc = self.RPN_Conv.weight.permute(0, 2, 3, 1).reshape(512, -1)            # tap-major [512, 9*1024]
            wh = torch.zeros(n_pad, 512, device=wc.device)
            wh[:2 * A] = self.RPN_cls_score.weight.flatten(1)
            wh[2 * A:6 * A] = self.RPN_bbox_pred.weight.flatten(1)
            bh = torch.zeros(n_pad, device=wc.device)
            bh[:2 * A] = self.RPN_cls_score.bias
            bh[2 * A:6 * A] = self.RPN_bbox_pred.bias
            keep = [mat(wc), self.RPN_Conv.bias.detach().float().contiguous(), mat(wh), bh.contiguous()]
        w = RPNWeights()
        w.dtype, w.round_tf32, w.A, w.n_pad = L.MODES[mode], 1 if mode == "tf32" else 0, A, n_pad
        w.conv.w, w.conv.bias, w.heads.w, w.heads.bias = (t.data_ptr() for t in keep)
        self._packed = (w, keep)
        return w

    def rpn_outputs(self, base_feat, im_info, want_reference_tensors=False):
        """-> proposals [B, H*W*A, 4], fg scores [B, H*W*A] (+ rpn_cls_prob [B,2A,H,W], rpn_bbox_pred [B,4A,H,W])."""
        lib = L.load()
        ops._need_cuda(base_feat, im_info)
        if base_feat.dtype != torch.float32 or base_feat.dim() != 4 or base_feat.shape[1] != self.din:
            raise RuntimeError("ait_b200._RPN: base_feat must be float32 [B, %d, H, W]" % self.din)
        fp = fingerprint(self)                  # in-place parameter updates since the last packing?
        if self._packed is None or self._packed_fp != fp:
            self._packed_fp = fp
            self._pack()
        w = self._packed[0]
        base_feat = base_feat.contiguous()
        B, _, H, W = base_feat.shape
        A, dev = w.A, base_feat.device
        props = torch.empty((B, H * W * A, 4), dtype=torch.float32, device=dev)
        fg = torch.empty((B, H * W * A), dtype=torch.float32, device=dev)
        cls_prob = torch.empty((B, 2 * A, H, W), dtype=torch.float32, device=dev) if want_reference_tensors else None
        bbox_pred = torch.empty((B, 4 * A, H, W), dtype=torch.float32, device=dev) if want_reference_tensors else None
        nbytes = lib.aitb_rpn_workspace_bytes(B, H, W, w.dtype)
        ws = ops._workspace(nbytes, dev, "rpn")
        with torch.cuda.device(dev):
            L.check(lib.aitb_rpn_forward(C.byref(w), L.ptr(base_feat), B, H, W, L.ptr(self._anchors),
                                         L.ptr(im_info.contiguous().float()), float(self.feat_stride), L.ptr(props),
                                         L.ptr(fg), L.ptr(cls_prob), L.ptr(bbox_pred), L.ptr(ws), nbytes, L.stream_ptr()))
        if want_reference_tensors:
            return props, fg, cls_prob, bbox_pred
        return props, fg

    def forward(self, base_feat, im_info, gt_boxes=None, num_boxes=None):
        """-> (rois [B, post_nms_topN, 5], rpn_loss_cls, rpn_loss_box) like rpn.py:110 (losses are 0 in eval)."""
        if self.training:
            return self._forward_train(base_feat, im_info, gt_boxes, num_boxes)
        c = self.cfg["TEST"]
        self.rpn_loss_cls = 0                    # rpn.py:92-93: reset on every forward
        self.rpn_loss_box = 0
        props, fg = self.rpn_outputs(base_feat, im_info)
        rois, _ = propose_rois(props, fg, c["pre_nms_topN"], c["post_nms_topN"], c["nms_thresh"])
        return rois, self.rpn_loss_cls, self.rpn_loss_box

    def _forward_train(self, base_feat, im_info, gt_boxes, num_boxes):
        """rpn.py:66-140 in .train(): -> (rois [B, TRAIN post_nms_topN, 5], rpn_loss_cls, rpn_loss_box); the losses carry the
        graph to base_feat and the six RPN parameters (fp32 storage / tf32 tensor-core math)."""
        from . import rpn_train, targets
        from .proposal import rpn_decode
        if gt_boxes is None:
            raise AssertionError("ait_b200._RPN: gt_boxes is required in .train() (rpn.py:98)")
        ops._need_cuda(base_feat, im_info, gt_boxes)
        score, bbox = rpn_train.rpn_head_train(self, base_feat)                  # rpn_cls_score, rpn_bbox_pred (:70-83)
        B, A2, H, W = score.shape
        with torch.no_grad():                                                    # the proposal layer sees `.data` (:89-90)
            prob = torch.softmax(score.view(B, 2, (A2 // 2) * H, W), dim=1).view(B, A2, H, W)     # :75-77
            props, fg = rpn_decode(prob, bbox.detach(), self._anchors, im_info, self.feat_stride)
            c = self.cfg["TRAIN"]
            rois, _ = propose_rois(props, fg, c["pre_nms_topN"], c["post_nms_topN"], c["nms_thresh"])
            if self._anchor_target is None:
                self._anchor_target = targets.AnchorTargetLayer(self.feat_stride, self.anchor_scales, self.anchor_ratios).to(
                    base_feat.device)
            rpn_data = self._anchor_target((score.detach(), gt_boxes, im_info, num_boxes))          # :100
        self.rpn_loss_cls, self.rpn_loss_box = targets.rpn_losses(score, bbox, rpn_data)            # :102-126
        return rois, self.rpn_loss_cls, self.rpn_loss_box
