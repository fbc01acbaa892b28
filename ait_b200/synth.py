"""Seeded synthetic inputs of the benchmark workload: ResNet-50 C4 maps of a
600x1000 image, 128x128 query features, RPN box/score outputs, random-init head weights."""
import torch

from .proposal import clip_boxes, decode_boxes, shifted_anchors

IM_H, IM_W = 600, 1000
FEAT_H, FEAT_W = 38, 63          # RCNN_base(1x3x600x1000) -> [1,1024,38,63] (Caffe-style strides)
VOC_SCALES, COCO_SCALES = (8, 16, 32), (4, 8, 16, 32)


def unit_seed(unit):
    return 1234 + int(unit)


def c4_map(unit, channels=1024, h=FEAT_H, w=FEAT_W):
    g = torch.Generator().manual_seed(unit_seed(unit))
    return torch.relu(torch.randn(channels, h, w, generator=g))


def query_feat(unit, channels=1024):
    g = torch.Generator().manual_seed(unit_seed(unit) + 500009)
    return torch.relu(torch.randn(channels, 8, 8, generator=g))


def rpn_outputs(unit, scales=VOC_SCALES, h=FEAT_H, w=FEAT_W):
    """(proposals [K*A,4], scores [K*A]): anchors + N(0, 0.2^2) deltas decoded and clipped to the image;
    scores = a random permutation of linspace(0, 1, K*A) (tie-free)."""
    g = torch.Generator().manual_seed(unit_seed(unit) + 1000003)
    anchors = shifted_anchors(h, w, scales=scales)
    n = anchors.shape[0]
    deltas = 0.2 * torch.randn(1, n, 4, generator=g)
    boxes = clip_boxes(decode_boxes(anchors, deltas), IM_H, IM_W)[0]
    scores = torch.linspace(0, 1, n)[torch.randperm(n, generator=g)]
    return boxes.contiguous(), scores.contiguous()


def random_rois(unit, n, batch_index=0):
    """isolated ROIAlign tests: x1~U(0,900), y1~U(0,500), w~U(16,600), h~U(16,400), clipped to the image."""
    g = torch.Generator().manual_seed(unit_seed(unit) + 2000003)
    x1 = torch.rand(n, generator=g) * 900
    y1 = torch.rand(n, generator=g) * 500
    bw = 16 + torch.rand(n, generator=g) * 584
    bh = 16 + torch.rand(n, generator=g) * 384
    x2 = (x1 + bw).clamp(max=IM_W - 1)
    y2 = (y1 + bh).clamp(max=IM_H - 1)
    return torch.stack([torch.full((n,), float(batch_index)), x1, y1, x2, y2], dim=1)


def make_head(seed=0, calibrated=False, randomize_bn=False, compute_dtype=torch.float32):
    """Random-init DetectionHead (reference init distributions).  `calibrated` raises the scale of the RCNN_cls_score /
    RCNN_bbox_pred weights above the stock init's degenerate 0.0057 +- 1e-6 scores -- but a random
    direction saturates instead (cls_prob 0.9991 .. 0.9999 on the benchmark inputs: measured), so a score comparison on
    it is still near-vacuous.  `spread_score_layer(head)` installs the score layer that does spread cls_prob (the one
    fitted for tests/golden/head_b2p4.pt); smoke(), bench.py and the parity tests use it.  `randomize_bn` gives the
    frozen BatchNorms non-trivial statistics."""
    from .head import DetectionHead
    torch.manual_seed(seed)
    head = DetectionHead(compute_dtype=compute_dtype)
    g = torch.Generator().manual_seed(seed + 77)
    if randomize_bn:
        for m in head.RCNN_top.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.copy_(0.1 * torch.randn(m.num_features, generator=g))
                m.running_var.copy_(0.5 + torch.rand(m.num_features, generator=g))
                m.weight.data.copy_(0.5 + torch.rand(m.num_features, generator=g))
                m.bias.data.copy_(0.1 * torch.randn(m.num_features, generator=g))
    if calibrated:
        with torch.no_grad():
            head.RCNN_cls_score[0].weight.copy_(torch.randn(8, 4096, generator=g) * 0.05)
            head.RCNN_cls_score[1].weight.copy_(torch.randn(2, 8, generator=g) * 1.0)
            head.RCNN_cls_score[1].bias.copy_(torch.randn(2, generator=g) * 0.1)
            head.RCNN_bbox_pred.weight.copy_(torch.randn(4, 2048, generator=g) * 0.01)
    return head.eval()


def spread_score_layer(head, path=None):
    """Install the spread-calibrated RCNN_cls_score of tests/golden/head_b2p4.pt (make_golden.py: first layer = the three
    principal directions of the golden pairs' features, unit spread) on a `make_head(seed=0, calibrated=True,
    randomize_bn=True)` head: cls_prob then spans 0.05 .. 0.99 on the smoke inputs and 0.03 .. 0.3 within one 300-proposal
    benchmark unit (different offsets per unit), so an absolute 1e-3 gate on it carries signal.  Returns the head."""
    import os
    if path is None:
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "head_b2p4.pt")
    state = torch.load(path, map_location="cpu", weights_only=False)["cls_score_state"]
    dev = head.RCNN_cls_score[0].weight.device
    head.RCNN_cls_score.load_state_dict({k: v.to(dev) for k, v in state.items()})
    if hasattr(head, "invalidate"):
        head.invalidate()
    return head
