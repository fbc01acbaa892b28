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
        if self._engine is None:
            self._engine = packing.HeadEngine(transformer=self.transformer, sk=self.sk, top=self.RCNN_top,
                                              cls_score=self.RCNN_cls_score, bbox_pred=self.RCNN_bbox_pred,
                                              dtype=self.compute_dtype)
        return self._engine

    def forward(self, non_img, non_qry, rois, taps=False):
        """non_img [B,1024,H,W] C4 map, non_qry [B,1024,8,8], rois [B,P,5] (batch idx, x1,y1,x2,y2)
        -> cls_prob [B,P,1], bbox_pred [B,P,4]   (+ dict of intermediates when taps=True)."""
        if self.training:
            raise RuntimeError("ait_b200.DetectionHead: inference only in this round (call .eval())")
        return self.engine().head_forward(non_img, non_qry, rois, taps=taps)
