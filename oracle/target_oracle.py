"""TEST INFRASTRUCTURE ONLY (never imported by the product).

CPU restatement of SURVEY section 8 row f4 -- the training-only samplers and losses around the head:

  bbox_overlaps_batch      lib/model/rpn/bbox_transform.py:167-257
  bbox_transform_batch     lib/model/rpn/bbox_transform.py:38-75
  anchor_target            lib/model/rpn/anchor_target_layer.py:49-199  (_AnchorTargetLayer.forward)
  proposal_target          lib/model/rpn/proposal_target_layer_cascade.py:33-220 (_ProposalTargetLayer.forward)
  smooth_l1                lib/model/utils/net_utils.py:75-89
  rpn_losses               lib/model/rpn/rpn.py:99-126
  rcnn_losses              lib/model/faster_rcnn/faster_rcnn_coatt_transformer_sk.py:340-361

Both samplers draw from numpy's GLOBAL generator exactly where the reference does (same calls, same
order, same arguments), so after `np.random.seed(s)` they consume the same stream and pick the same
anchors / rois.  Pinned to the unmodified reference by tests/golden/make_golden_targets.py ->
tests/golden/targets.pt and, where /root/reference exists, live (tests/test_oracle_pins.py).
"""
import numpy as np
import torch
import torch.nn.functional as F

# lib/model/utils/config.py:81-158 (cfg.TRAIN.*) and :23 (MARGIN)
TRAIN = dict(BATCH_SIZE=128, FG_FRACTION=0.25, FG_THRESH=0.5, BG_THRESH_HI=0.5, BG_THRESH_LO=0.0,
             BBOX_NORMALIZE_MEANS=(0.0, 0.0, 0.0, 0.0), BBOX_NORMALIZE_STDS=(0.1, 0.1, 0.2, 0.2),
             BBOX_INSIDE_WEIGHTS=(1.0, 1.0, 1.0, 1.0),
             RPN_POSITIVE_OVERLAP=0.7, RPN_NEGATIVE_OVERLAP=0.3, RPN_CLOBBER_POSITIVES=False, RPN_FG_FRACTION=0.5,
             RPN_BATCHSIZE=256, RPN_BBOX_INSIDE_WEIGHTS=(1.0, 1.0, 1.0, 1.0), RPN_POSITIVE_WEIGHT=-1.0, MARGIN=-0.3)


def bbox_overlaps_batch(anchors, gt_boxes):
    """anchors [N,4] or [B,N,4|5], gt_boxes [B,K,5] -> overlaps [B,N,K] (legacy +1 areas); an all-zero gt box
    (1x1 after the +1) gives overlap 0, a 1x1 anchor gives -1  (bbox_transform.py:167-257)."""
    B = gt_boxes.size(0)
    if anchors.dim() == 2:
        anchors = anchors.view(1, -1, 4).expand(B, -1, 4)
    elif anchors.size(2) == 5:
        anchors = anchors[:, :, 1:5]
    gt = gt_boxes[:, :, :4]
    gx = gt[:, :, 2] - gt[:, :, 0] + 1
    gy = gt[:, :, 3] - gt[:, :, 1] + 1
    g_area = (gx * gy).unsqueeze(1)
    ax = anchors[:, :, 2] - anchors[:, :, 0] + 1
    ay = anchors[:, :, 3] - anchors[:, :, 1] + 1
    a_area = (ax * ay).unsqueeze(2)
    g_zero = (gx == 1) & (gy == 1)
    a_zero = (ax == 1) & (ay == 1)
    a = anchors.unsqueeze(2)
    g = gt.unsqueeze(1)
    iw = (torch.min(a[..., 2], g[..., 2]) - torch.max(a[..., 0], g[..., 0]) + 1).clamp(min=0)
    ih = (torch.min(a[..., 3], g[..., 3]) - torch.max(a[..., 1], g[..., 1]) + 1).clamp(min=0)
    ua = a_area + g_area - iw * ih
    ov = iw * ih / ua
    ov = ov.masked_fill(g_zero.unsqueeze(1), 0)
    ov = ov.masked_fill(a_zero.unsqueeze(2), -1)
    return ov


def bbox_transform_batch(ex, gt):
    """ex [N,4] or [B,N,4], gt [B,N,4] -> (dx, dy, dw, dh) [B,N,4]  (bbox_transform.py:38-75)."""
    if ex.dim() == 2:
        ex = ex.unsqueeze(0)
    ew = ex[..., 2] - ex[..., 0] + 1.0
    eh = ex[..., 3] - ex[..., 1] + 1.0
    ecx = ex[..., 0] + 0.5 * ew
    ecy = ex[..., 1] + 0.5 * eh
    gw = gt[..., 2] - gt[..., 0] + 1.0
    gh = gt[..., 3] - gt[..., 1] + 1.0
    gcx = gt[..., 0] + 0.5 * gw
    gcy = gt[..., 1] + 0.5 * gh
    return torch.stack(((gcx - ecx) / ew, (gcy - ecy) / eh, torch.log(gw / ew), torch.log(gh / eh)), 2)


def all_anchors(base_anchors, H, W, feat_stride):
    """[H*W*A, 4]: cell-major, A anchors per cell (anchor_target_layer.py:67-81)."""
    sx = torch.arange(W, dtype=torch.float32) * feat_stride
    sy = torch.arange(H, dtype=torch.float32) * feat_stride
    yy, xx = torch.meshgrid(sy, sx, indexing="ij")
    shifts = torch.stack([xx.reshape(-1), yy.reshape(-1), xx.reshape(-1), yy.reshape(-1)], 1)
    return (base_anchors.view(1, -1, 4).float() + shifts.view(-1, 1, 4)).reshape(-1, 4)


def anchor_target(base_anchors, H, W, feat_stride, gt_boxes, im_info, cfg=TRAIN, stage=None):
    """_AnchorTargetLayer.forward (anchor_target_layer.py:49-199).  Returns the four reference outputs
    labels [B,1,A*H,W], bbox_targets / inside / outside weights [B,4A,H,W].  `stage` (a dict) receives the
    pre-sampling labels for the deterministic parity checks."""
    B = gt_boxes.size(0)
    A = base_anchors.size(0)
    anchors_all = all_anchors(base_anchors, H, W, feat_stride)
    total = anchors_all.size(0)
    keep = ((anchors_all[:, 0] >= 0) & (anchors_all[:, 1] >= 0) &
            (anchors_all[:, 2] < int(im_info[0][1])) & (anchors_all[:, 3] < int(im_info[0][0])))   # :85-88
    inds = torch.nonzero(keep).view(-1)
    anchors = anchors_all[inds]
    n_in = inds.numel()
    labels = gt_boxes.new_full((B, n_in), -1)
    ov = bbox_overlaps_batch(anchors, gt_boxes)
    max_ov, argmax = torch.max(ov, 2)
    gt_max, _ = torch.max(ov, 1)
    labels[max_ov < cfg["RPN_NEGATIVE_OVERLAP"]] = 0                       # :110-111 (no clobber)
    gt_max[gt_max == 0] = 1e-5                                            # :113
    hit = torch.sum(ov.eq(gt_max.view(B, 1, -1).expand_as(ov)), 2)        # :114
    if torch.sum(hit) > 0:
        labels[hit > 0] = 1
    labels[max_ov >= cfg["RPN_POSITIVE_OVERLAP"]] = 1                      # :120
    if stage is not None:
        stage["labels_presample"] = labels.clone()
        stage["inds_inside"] = inds.clone()
    num_fg = int(cfg["RPN_FG_FRACTION"] * cfg["RPN_BATCHSIZE"])
    sum_fg = torch.sum((labels == 1).int(), 1)
    sum_bg = torch.sum((labels == 0).int(), 1)
    for i in range(B):                                                     # :130-152
        if sum_fg[i] > num_fg:
            fg_inds = torch.nonzero(labels[i] == 1).view(-1)
            rand = torch.from_numpy(np.random.permutation(fg_inds.size(0))).long()
            labels[i][fg_inds[rand[:fg_inds.size(0) - num_fg]]] = -1
        num_bg = cfg["RPN_BATCHSIZE"] - torch.sum((labels == 1).int(), 1)[i]
        if sum_bg[i] > num_bg:
            bg_inds = torch.nonzero(labels[i] == 0).view(-1)
            rand = torch.from_numpy(np.random.permutation(bg_inds.size(0))).long()
            labels[i][bg_inds[rand[:bg_inds.size(0) - num_bg]]] = -1
    gt_sel = torch.gather(gt_boxes[:, :, :4], 1, argmax.unsqueeze(2).expand(B, n_in, 4))
    targets = bbox_transform_batch(anchors, gt_sel)                        # :157
    inside = gt_boxes.new_zeros((B, n_in))
    outside = gt_boxes.new_zeros((B, n_in))
    inside[labels == 1] = cfg["RPN_BBOX_INSIDE_WEIGHTS"][0]
    # uniform weights from the example count of the LAST image (`labels[i]` after the loop, :163)
    num_examples = torch.sum(labels[B - 1] >= 0)
    w = 1.0 / num_examples.item()
    outside[labels == 1] = w
    outside[labels == 0] = w

    def unmap(data, fill):                                                 # :201-211
        shape = (B, total) + tuple(data.shape[2:])
        ret = gt_boxes.new_full(shape, fill)
        ret[:, inds] = data
        return ret

    labels = unmap(labels, -1).view(B, H, W, A).permute(0, 3, 1, 2).contiguous().view(B, 1, A * H, W)
    targets = unmap(targets, 0).view(B, H, W, A * 4).permute(0, 3, 1, 2).contiguous()
    inside = unmap(inside, 0).view(B, total, 1).expand(B, total, 4).contiguous().view(B, H, W, 4 * A).permute(0, 3, 1, 2).contiguous()
    outside = unmap(outside, 0).view(B, total, 1).expand(B, total, 4).contiguous().view(B, H, W, 4 * A).permute(0, 3, 1, 2).contiguous()
    return labels, targets, inside, outside


def proposal_target(all_rois, gt_boxes, cfg=TRAIN, stage=None):
    """_ProposalTargetLayer.forward (proposal_target_layer_cascade.py:33-220): rois [B,R,5], gt_boxes [B,K,5]
    -> rois [B,128,5], labels [B,128], bbox_targets, inside weights, outside weights [B,128,4]."""
    B = gt_boxes.size(0)
    gt_append = gt_boxes.new_zeros(gt_boxes.size())
    gt_append[:, :, 1:5] = gt_boxes[:, :, :4]
    all_rois = torch.cat([all_rois, gt_append], 1)                          # :45
    rois_per_image = int(cfg["BATCH_SIZE"] / 1)
    fg_per_image = int(np.round(cfg["FG_FRACTION"] * rois_per_image)) or 1
    ov = bbox_overlaps_batch(all_rois, gt_boxes)
    max_ov, assign = torch.max(ov, 2)
    labels = torch.gather(gt_boxes[:, :, 4], 1, assign)                    # :143-147
    if stage is not None:
        stage["max_overlaps"], stage["assignment"] = max_ov.clone(), assign.clone()
    labels_b = labels.new_zeros((B, rois_per_image))
    rois_b = all_rois.new_zeros((B, rois_per_image, 5))
    gt_rois_b = all_rois.new_zeros((B, rois_per_image, 5))
    for i in range(B):                                                     # :154-214
        fg_inds = torch.nonzero(max_ov[i] >= cfg["FG_THRESH"]).view(-1)
        bg_inds = torch.nonzero((max_ov[i] < cfg["BG_THRESH_HI"]) & (max_ov[i] >= cfg["BG_THRESH_LO"])).view(-1)
        nf, nb = fg_inds.numel(), bg_inds.numel()
        if nf > 0 and nb > 0:
            fg_this = min(fg_per_image, nf)
            rand = torch.from_numpy(np.random.permutation(nf)).long()
            fg_inds = fg_inds[rand[:fg_this]]
            bg_this = rois_per_image - fg_this
            rand = torch.from_numpy(np.floor(np.random.rand(bg_this) * nb)).long()
            bg_inds = bg_inds[rand]
        elif nf > 0:
            rand = torch.from_numpy(np.floor(np.random.rand(rois_per_image) * nf)).long()
            fg_inds = fg_inds[rand]
            fg_this = rois_per_image
        elif nb > 0:
            rand = torch.from_numpy(np.floor(np.random.rand(rois_per_image) * nb)).long()
            bg_inds = bg_inds[rand]
            fg_this = 0
        else:
            raise ValueError("bg_num_rois = 0 and fg_num_rois = 0, this should not happen!")
        keep = torch.cat([fg_inds, bg_inds], 0)
        labels_b[i].copy_(labels[i][keep])
        if fg_this < rois_per_image:
            labels_b[i][fg_this:] = 0
        rois_b[i] = all_rois[i][keep]
        rois_b[i, :, 0] = i
        gt_rois_b[i] = gt_boxes[i][assign[i][keep]]
    t = bbox_transform_batch(rois_b[:, :, 1:5], gt_rois_b[:, :, :4])
    means = t.new_tensor(cfg["BBOX_NORMALIZE_MEANS"])
    stds = t.new_tensor(cfg["BBOX_NORMALIZE_STDS"])
    t = (t - means.expand_as(t)) / stds.expand_as(t)                        # :118-121
    fgm = (labels_b > 0).unsqueeze(2).expand_as(t)
    targets = torch.where(fgm, t, torch.zeros_like(t))                      # :92-103
    inside = torch.where(fgm, t.new_tensor(cfg["BBOX_INSIDE_WEIGHTS"]).expand_as(t), torch.zeros_like(t))
    outside = (inside > 0).float()
    return rois_b, labels_b, targets, inside, outside


def smooth_l1(pred, targets, inside, outside, sigma=1.0, dim=(1,)):
    """_smooth_l1_loss (net_utils.py:75-89)."""
    s2 = sigma ** 2
    d = inside * (pred - targets)
    ad = d.abs()
    sign = (ad < 1.0 / s2).detach().float()
    loss = outside * (d.pow(2) * (s2 / 2.0) * sign + (ad - 0.5 / s2) * (1.0 - sign))
    for i in sorted(dim, reverse=True):
        loss = loss.sum(i)
    return loss.mean()


def rpn_losses(rpn_cls_score, rpn_bbox_pred, labels, targets, inside, outside):
    """rpn.py:99-126.  rpn_cls_score [B,2A,H,W] (bg channels first), rpn_bbox_pred [B,4A,H,W], anchor-target outputs
    -> (rpn_loss_cls, rpn_loss_box)."""
    B = rpn_cls_score.size(0)
    score = rpn_cls_score.view(B, 2, -1, rpn_cls_score.size(3))             # rpn.py:70 reshape(x, 2)
    score = score.permute(0, 2, 3, 1).contiguous().view(B, -1, 2)
    lab = labels.view(B, -1)
    keep = lab.view(-1).ne(-1).nonzero().view(-1)
    loss_cls = F.cross_entropy(score.view(-1, 2).index_select(0, keep), lab.view(-1).index_select(0, keep).long())
    loss_box = smooth_l1(rpn_bbox_pred, targets, inside, outside, sigma=3, dim=(1, 2, 3))
    return loss_cls, loss_box


def rcnn_losses(score, bbox_pred, rois_label, rois_target, rois_inside, rois_outside, bs, margin=TRAIN["MARGIN"]):
    """faster_rcnn_coatt_transformer_sk.py:340-361.  score [bs*P,2], bbox_pred [bs*P,4], rois_label [bs*P] ->
    (RCNN_loss_cls, margin_loss, RCNN_loss_bbox)."""
    prob = F.softmax(score, 1)[:, 1]
    lab = rois_label.view(bs, -1).float()
    gt_map = torch.abs(lab.unsqueeze(1) - lab.unsqueeze(-1))
    p = prob.view(bs, -1)
    pr_map = torch.abs(p.unsqueeze(1) - p.unsqueeze(-1))
    target = -((gt_map - 1) ** 2) + gt_map
    loss_cls = F.cross_entropy(score, rois_label.view(-1).long())
    margin_loss = 3 * F.margin_ranking_loss(pr_map, gt_map, target, margin=margin)
    loss_bbox = smooth_l1(bbox_pred, rois_target.view(-1, 4), rois_inside.view(-1, 4), rois_outside.view(-1, 4))
    return loss_cls, margin_loss, loss_bbox


# ---- seeded synthetic inputs shared by tests/golden/make_golden_targets.py and the tests ----
def synth_gt_boxes(seed, B, K=20, im_h=300.0, im_w=500.0, n_min=2, n_max=6):
    """gt_boxes [B,K,5] (x1,y1,x2,y2,cls=1), zero rows after the image's boxes; num_boxes [B]."""
    g = torch.Generator().manual_seed(seed)
    gt = torch.zeros(B, K, 5)
    nb = torch.zeros(B, dtype=torch.long)
    for b in range(B):
        n = n_min + (b * 3) % (n_max - n_min + 1)
        x1 = torch.rand(n, generator=g) * (im_w * 0.6)
        y1 = torch.rand(n, generator=g) * (im_h * 0.6)
        w = 30 + torch.rand(n, generator=g) * (im_w * 0.35)
        h = 30 + torch.rand(n, generator=g) * (im_h * 0.35)
        gt[b, :n, 0], gt[b, :n, 1] = x1, y1
        gt[b, :n, 2], gt[b, :n, 3] = (x1 + w).clamp(max=im_w - 1), (y1 + h).clamp(max=im_h - 1)
        gt[b, :n, 4] = 1
        nb[b] = n
    return gt, nb


def synth_rois(seed, B, R, gt_boxes, im_h=300.0, im_w=500.0):
    """rois [B,R,5]: a third jittered copies of gt boxes (foreground candidates), the rest uniform boxes; a few
    all-zero rows at the end like the proposal layer's padding."""
    g = torch.Generator().manual_seed(seed)
    rois = torch.zeros(B, R, 5)
    for b in range(B):
        n_gt = int((gt_boxes[b, :, 4] > 0).sum())
        xy = torch.rand(R, 2, generator=g) * torch.tensor([im_w * 0.8, im_h * 0.75])
        wh = 16 + torch.rand(R, 2, generator=g) * torch.tensor([im_w * 0.4, im_h * 0.45])
        box = torch.cat([xy, xy + wh], 1)
        nj = R // 3
        src = gt_boxes[b, torch.randint(0, max(n_gt, 1), (nj,), generator=g), :4]
        box[:nj] = src + torch.randn(nj, 4, generator=g) * 12.0
        box[:, 0::2] = box[:, 0::2].clamp(0, im_w - 1)
        box[:, 1::2] = box[:, 1::2].clamp(0, im_h - 1)
        box[:, 2] = torch.maximum(box[:, 2], box[:, 0])
        box[:, 3] = torch.maximum(box[:, 3], box[:, 1])
        rois[b, :, 0] = b
        rois[b, :, 1:] = box
        rois[b, R - 5:, :] = 0
    return rois
