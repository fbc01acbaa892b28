"""Everything of `_fasterRCNN.forward` after the backbone
(lib/model/faster_rcnn/faster_rcnn_coatt_transformer_sk.py:229-337): co-attention -> RPN head -> proposal layer ->
ROIAlign -> AIT -> SKNet -> RCNN_top -> score / box heads, plus (optionally) the detection post-processing of
test_net_voc.py:380-446 -- `.eval()`: `forward`; `.train()`: `training_step` (the reference's training forward, :229-361:
+ anchor targets, proposal targets and the five losses, every stage differentiable on the device).  Sub-module names are the detector's (coattention_module, RCNN_rpn, RCNN_roi_align,
transformer, sk, RCNN_top, RCNN_cls_score, RCNN_bbox_pred), so `load_state_dict(detector.state_dict(), strict=False)`
picks up a reference checkpoint.  Every stage runs in libaitb200; no tensor leaves the device between stages.
"""
import torch
import torch.nn as nn

from .coattention import CoAttentionModule
from .head import DetectionHead
from .proposal import detections
from .rpn import _RPN


class DetectorTail(DetectionHead):
    def __init__(self, channels=1024, compute_dtype=torch.float32, rpn_cfg=None):
        super().__init__(channels=channels, compute_dtype=compute_dtype)
        self.coattention_module = CoAttentionModule(channels)
        self.RCNN_rpn = _RPN(channels, cfg=rpn_cfg, compute_dtype=compute_dtype)

    def forward(self, image_feat, query_feat, im_info, postprocess=False, **post_kw):
        """image_feat [B,1024,H,W], query_feat [B,1024,8,8] (RCNN_base outputs), im_info [B,3]
        -> rois [B,P,5], cls_prob [B,P,1], bbox_pred [B,P,4]   (+ dets [B,P,5], n_det [B] with postprocess=True)."""
        if self.training:
            raise RuntimeError("ait_b200.DetectorTail: .forward is the inference path (call .eval()); the training forward is "
                               ".training_step(image_feat, query_feat, im_info, gt_boxes, num_boxes)")
        non_img, non_qry = self.coattention_module(image_feat, query_feat)
        rois, _, _ = self.RCNN_rpn(non_img, im_info, None, None)
        cls_prob, bbox_pred = self.engine().head_forward(non_img, non_qry, rois)
        if not postprocess:
            return rois, cls_prob, bbox_pred
        dets, n_det = detections(rois, cls_prob, bbox_pred, im_info, **post_kw)
        return rois, cls_prob, bbox_pred, dets, n_det

    def training_step(self, image_feat, query_feat, im_info, gt_boxes, num_boxes, sampler=None):
        """The reference's training forward after the backbone (faster_rcnn_coatt_transformer_sk.py:229-361):
        co-attention -> `_RPN` (head, proposal layer with the TRAIN settings, anchor targets, RPN losses) ->
        `ProposalTargetLayer` -> ROIAlign -> AIT -> SKNet -> RCNN_top -> heads -> RCNN losses.
        -> (rois [B,R,5], rpn_loss_cls, rpn_loss_bbox, RCNN_loss_cls, margin_loss, RCNN_loss_bbox, rois_label [B*R]); the five
        losses carry the graph to image_feat, query_feat and every trainable parameter.  sampler: a `ProposalTargetLayer`
        (default: one per module with the reference's numpy sampling, seeded by np.random.seed)."""
        from . import targets
        if not self.training:
            raise RuntimeError("ait_b200.DetectorTail.training_step: call .train() first")
        non_img, non_qry = self.coattention_module(image_feat, query_feat)                                   # :236
        rois, rpn_loss_cls, rpn_loss_bbox = self.RCNN_rpn(non_img, im_info, gt_boxes, num_boxes)             # :249
        if sampler is None:
            if getattr(self, "_proposal_target", None) is None:
                self._proposal_target = targets.ProposalTargetLayer(2)
            sampler = self._proposal_target
        with torch.no_grad():
            rois, label, tgt, inw, outw = sampler(rois, gt_boxes, num_boxes)                                 # :257-258
        label = label.view(-1).long()
        loss_cls, margin_loss, loss_bbox = self.training_losses(non_img, non_qry, rois, label, tgt, inw, outw)   # :273-361
        return rois, rpn_loss_cls, rpn_loss_bbox, loss_cls, margin_loss, loss_bbox, label
