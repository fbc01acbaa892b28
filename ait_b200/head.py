"""The per-(image, query) detection head as one module: the slice of `_fasterRCNN.forward` from the
ROIAlign call to cls_prob / bbox_pred
(lib/model/faster_rcnn/faster_rcnn_coatt_transformer_sk.py:273-337), with the sub-module names of
`_fasterRCNN` / `resnet` so that a reference detector checkpoint loads with strict=False:

    RCNN_roi_align, transformer.*, sk.*, RCNN_top.0.*, RCNN_cls_score.{0,1}.*, RCNN_bbox_pred.*
"""
import torch
import torch.nn as nn

from . import packing
from .modules import SKNet, make_layer4
from .roi_layers import ROIAlign
from ._lib import mode_name as packing_mode
from .system.Models import Transformer

POOLING_SIZE = 7          # cfg.POOLING_SIZE (lib/model/utils/config.py:294)
FEAT_STRIDE = 16          # cfg.FEAT_STRIDE


class DetectionHead(nn.Module):
    def __init__(self, channels=1024, class_agnostic=True, compute_dtype=torch.float32, dropout=0.1):
        super().__init__()
        if channels != 1024 or not class_agnostic:
            raise NotImplementedError("ait_b200.DetectionHead: ResNet-50 C4 (1024 ch), class-agnostic boxes only")
        self.compute_dtype = compute_dtype
        self.RCNN_roi_align = ROIAlign((POOLING_SIZE, POOLING_SIZE), 1.0 / FEAT_STRIDE, 0)
        self.sk = SKNet(channels=channels, compute_dtype=compute_dtype)
        self.transformer = Transformer(d_k=64, d_v=64, d_model=channels // 2, d_word_vec=channels // 2,
                                       d_inner=channels * 2, n_position=8 * 8, n_layers=1, n_head=8,
                                       dropout=dropout, compute_dtype=compute_dtype)
        self.RCNN_top = nn.Sequential(make_layer4())
        self.RCNN_cls_score = nn.Sequential(nn.Linear(2048 * 2, 8), nn.Linear(8, 2))
        self.RCNN_bbox_pred = nn.Linear(2048, 4)
        # _init_weights (faster_rcnn_coatt_transformer_sk.py:389-394)
        for m, std in ((self.RCNN_cls_score[0], 0.01), (self.RCNN_cls_score[1], 0.01), (self.RCNN_bbox_pred, 0.001)):
            m.weight.data.normal_(0, std)
            m.bias.data.zero_()
        self._engine = None

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def _load_from_state_dict(self, *a, **k):
        super()._load_from_state_dict(*a, **k)
        self._engine = None

    def invalidate(self):
        self._engine = None

    def engine(self):
        fp = packing.fingerprint(self)          # in-place parameter updates since the last packing?
        if self._engine is None or self._engine_fp != fp:
            self._engine_fp = fp
            self._engine = packing.HeadEngine(transformer=self.transformer, sk=self.sk, top=self.RCNN_top,
                                              cls_score=self.RCNN_cls_score, bbox_pred=self.RCNN_bbox_pred,
                                              dtype=self.compute_dtype)
        return self._engine

    def forward(self, non_img, non_qry, rois, taps=False):
        """non_img [B,1024,H,W] C4 map, non_qry [B,1024,8,8], rois [B,P,5] (batch idx, x1,y1,x2,y2)
        -> cls_prob [B,P,1], bbox_pred [B,P,4]   (+ dict of intermediates when taps=True)."""
        if self.training:
            raise RuntimeError("ait_b200.DetectionHead: .forward is the inference engine (call .eval()); the training step "
                               "is .forward_train / .training_losses")
        return self.engine().head_forward(non_img, non_qry, rois, taps=taps)

    def forward_train(self, non_img, non_qry, rois):
        """The same slice of `_fasterRCNN.forward` (:273-335) as a differentiable training step (BASELINE config 4):
        ROIAlign -> AIT -> SKNet -> `_head_to_tail` (pairs and queries) -> bbox / score heads, every stage with the
        library's own forward-keeping-activations + backward (`_ROIAlign`, `_AITTrainFunction`, sk_train, top_train,
        targets.score_heads; fp32 storage, tf32 tensor-core math; the Transformer applies its training-mode dropout, see system/Models.py).  Returns (score [B*P,2] logits,
        bbox_pred [B*P,4]) -- what the reference's losses consume (:349-361).  Gradients reach non_img, non_qry and
        every trainable parameter (AIT 46, the four SK convolutions + biases, the ten layer-4 convolutions -- BatchNorm is
        frozen like `set_bn_fix` --, the three Linear layers); `sk.*.fc` / `sk.*.sk` get none, as in the reference,
        whose SKBlock.forward discards them."""
        from . import sk_train, targets, top_train
        if not self.training:
            raise RuntimeError("ait_b200.DetectionHead.forward_train: call .train() first")
        if rois.dim() != 3 or rois.shape[2] != 5 or rois.shape[0] != non_qry.shape[0]:
            raise RuntimeError("forward_train: rois must be [B,P,5] with B = number of (image, query) units")
        P = rois.shape[1]
        props = self.RCNN_roi_align(non_img, rois.reshape(-1, 5))                 # :279
        # stage-to-stage hand-over in the GEMM operand layout (token-major / channels-last, tf32-rounded): AIT -> SKNet -> layer4
        # without an NCHW round trip, forward or backward
        # (the bf16 AIT training configuration returns fp32 NCHW: the rest of the head trains in fp32 storage / tf32 math)
        tm = packing_mode(self.transformer.compute_dtype) != "bf16"
        props = self.transformer(x_props=props, x_query=non_qry, token_major_out=tm)           # :289
        props, query = sk_train.sknet_train(self.sk, props, non_qry, channels_last_out=True, channels_last_in=tm)   # :294
        pf = top_train.head_to_tail_train(self.RCNN_top, props, channels_last=True)            # :299
        qf = top_train.head_to_tail_train(self.RCNN_top, query, channels_last=True)            # :300
        return targets.score_heads(pf, qf, P, self.RCNN_bbox_pred, self.RCNN_cls_score)   # :318-335

    def training_losses(self, non_img, non_qry, rois, rois_label, rois_target, rois_inside_ws, rois_outside_ws):
        """(RCNN_loss_cls, margin_loss, RCNN_loss_bbox) of faster_rcnn_coatt_transformer_sk.py:340-361 for the units'
        sampled rois (the outputs of `ProposalTargetLayer`)."""
        from . import targets
        score, bbox_pred = self.forward_train(non_img, non_qry, rois)
        return targets.rcnn_losses(score, bbox_pred, rois_label, rois_target, rois_inside_ws, rois_outside_ws,
                                   rois.shape[0])
