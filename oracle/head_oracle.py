"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch functional, fp32 or fp64) of the reference's
floating-point head path.  Never imported by ait_b200/.  Each function cites the reference lines it
follows; `tests/test_oracle_pins.py` pins all of them against the reference's own modules (imported
from /root/reference where present) and against the committed golden vectors.

The arithmetic is the reference's: the same ATen CPU kernels (F.conv2d, F.linear, matmul, softmax,
F.layer_norm) in the same order, with the literal per-proposal recomputation of the decoder
self-attention that the reference performs (system/Models.py:250-253).

Weights are passed as a flat dict with the reference's state_dict keys, prefixed like
`_fasterRCNN`'s sub-modules: transformer.*, sk.*, RCNN_top.0.*, RCNN_cls_score.*, RCNN_bbox_pred.*.
"""
import torch
import torch.nn.functional as F

from . import c_ops


# Tests only: callable(tag, pre_activation) -> multiplier mask (or None = plain ReLU).  The training parity tests rebuild the
# fp64 reference graph with the DEVICE's ReLU decisions (a tf32 forward flips the mask of hidden units whose pre-activation is
# within rounding error of zero; with the masks injected both sides differentiate the same piecewise-linear function).
# Tags: "enc_ffn", "dec_ffn", "l4.<call index>.<block>.<1|2|3>" (call index: head_to_tail calls in order, props then query).
RELU_HOOK = None
_l4_calls = [0]


def _relu(x, tag):
    if RELU_HOOK is not None:
        m = RELU_HOOK(tag, x)
        if m is not None:
            return x * m.to(device=x.device, dtype=x.dtype)
    return F.relu(x)


def _sub(sd, prefix):
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


def _cast(sd, dtype):
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}


# ------------------------------------------------------------------------------------------------
# AIT  (lib/model/system/Models.py, Layers.py, SubLayers.py, Modules.py)
# ------------------------------------------------------------------------------------------------
def _attention(q, k, v, mask, drop=None):
    """ScaledDotProductAttention.forward (system/Modules.py:16-29), temperature = sqrt(64) = 8.
    drop: the multipliers (0 | 1 / (1 - p)) nn.Dropout would apply to the probabilities in .train() (:24), injected."""
    attn = torch.matmul(q / 8.0, k.transpose(2, 3))
    attn = attn.masked_fill(mask == 0, -1e9)
    attn = F.softmax(attn, dim=-1)
    if drop is not None:
        attn = attn * drop
    return torch.matmul(attn, v)


def _sh_block(x, w):
    """SHBlock.forward (system/SubLayers.py:22-39): x [b, 8, T, 64]."""
    b, n_head, T, Cc = x.shape
    u = x.sum(dim=1)
    s = u.transpose(1, 2).mean(dim=2)                      # AdaptiveAvgPool1d(1) over T
    g = F.linear(s, w["sh.sk.weight"], w["sh.sk.bias"]).view(b, n_head, Cc)
    g = F.softmax(g, dim=1).unsqueeze(2)
    return x * g


def _mha(w, q_in, k_in, v_in, mask, drop_attn=None, drop_fc=None):
    """MultiHeadAttention.forward (system/SubLayers.py:68-102), n_head = 8, d_k = d_v = 64; eval mode unless the dropout
    multipliers of .train() are injected (drop_attn [b,8,lq,lk] at Modules.py:24, drop_fc [b,lq,512] at SubLayers.py:97)."""
    b, lq, lk = q_in.size(0), q_in.size(1), k_in.size(1)
    residual = q_in
    q = F.linear(q_in, w["w_qs.weight"]).view(b, lq, 8, 64).transpose(1, 2)
    k = F.linear(k_in, w["w_ks.weight"]).view(b, lk, 8, 64).transpose(1, 2)
    v = F.linear(v_in, w["w_vs.weight"]).view(b, lk, 8, 64).transpose(1, 2)
    o = _attention(q, k, v, mask.unsqueeze(1), drop_attn)
    o = _sh_block(o, w).sum(dim=1, keepdim=True)            # selective heads, summed (:89-92)
    o = o.transpose(1, 2).contiguous().view(b, lq, -1)
    o = F.linear(o, w["fc.weight"])
    if drop_fc is not None:
        o = o * drop_fc
    o = o + residual
    return F.layer_norm(o, (512,), w["layer_norm.weight"], w["layer_norm.bias"], eps=1e-6)


def _ffn(w, x, drop=None, tag="ffn"):
    """PositionwiseFeedForward.forward (system/SubLayers.py:177-187); drop [b,l,512]: injected multipliers of :182."""
    y = F.linear(_relu(F.linear(x, w["w_1.weight"], w["w_1.bias"]), tag), w["w_2.weight"], w["w_2.bias"])
    if drop is not None:
        y = y * drop
    y = y + x
    return F.layer_norm(y, (512,), w["layer_norm.weight"], w["layer_norm.bias"], eps=1e-6)


def ait_forward(sd, x_props, x_query, dtype=torch.float32, return_enc=False, drop=None):
    """Transformer.forward (system/Models.py:231-280).  sd: the 48 transformer keys (no prefix).
    drop: .train() with the dropout multipliers INJECTED (the reference draws them from torch's generator; the device
    path from its own counter-based generator, `HeadEngine.dropout_masks`): dict of enc_emb / enc_slf_fc / dec_enc_fc /
    enc_ffn / dec_ffn [bp*64, 512], dec_emb / dec_slf_fc [bs*64, 512], enc_slf_attn / dec_enc_attn [bp,8,64,64],
    dec_slf_attn [bs,8,64,64].  The decoder-side masks are per UNIT and repeated over the unit's proposals (the device path
    runs the query side once per unit).  None = eval mode."""
    w = _cast(sd, dtype)
    dm = {} if drop is None else {k: v.to(device=x_props.device, dtype=dtype) for k, v in drop.items()}

    def rows(name, per_unit=False):          # [n*64, 512] -> [bp, 64, 512]
        if name not in dm:
            return None
        m = dm[name].view(-1, 64, 512)
        return m.repeat_interleave(num_props, dim=0) if per_unit else m

    def heads(name, per_unit=False):         # [n, 8, 64, 64] -> [bp, 8, 64, 64]
        if name not in dm:
            return None
        return dm[name].repeat_interleave(num_props, dim=0) if per_unit else dm[name]
    x_props, x_query = x_props.to(dtype), x_query.to(dtype)
    bp, bs = x_props.size(0), x_query.size(0)
    num_props = bp // bs
    xp = F.conv2d(x_props, w["enc_emb.0.weight"], w["enc_emb.0.bias"])           # :246
    xq = F.conv2d(x_query, w["dec_emb.0.weight"], w["dec_emb.0.bias"])           # :247
    rq = xq.unsqueeze(1).repeat(1, num_props, 1, 1, 1)                            # :250
    src = xp.view(bp, 512, -1).permute(0, 2, 1)                                   # [bp, 49, 512]
    trg = rq.view(bp, 512, -1).permute(0, 2, 1)                                   # [bp, 64, 512]
    n_s, n_t = src.size(1), trg.size(1)
    dev = x_props.device   # the reference builds masks / padding on the CPU and moves them (.to(device), :258-269)
    src_mask = torch.cat([torch.ones(bp, 1, n_s, dtype=torch.uint8, device=dev),
                          torch.zeros(bp, 1, n_t - n_s, dtype=torch.uint8, device=dev)], dim=2)            # :258-260
    trg_mask = (1 - torch.triu(torch.ones(1, n_t, n_t, device=dev), diagonal=1)).to(torch.uint8).expand(bp, n_t, n_t)  # :262-263
    src = torch.cat([src, torch.zeros(bp, n_t - n_s, 512, dtype=dtype, device=dev)], dim=1)   # :268-270
    # Encoder.forward (:83-111)
    e = src + w["encoder.position_enc.pos_table"][:, :n_t]
    if rows("enc_emb") is not None:                                                # Encoder.dropout (:98)
        e = e * rows("enc_emb")
    e = F.layer_norm(e, (512,), w["encoder.layer_norm.weight"], w["encoder.layer_norm.bias"], eps=1e-6)
    el = _sub(w, "encoder.layer_stack.0.")
    e = _mha(_sub(el, "slf_attn."), e, e, e, src_mask, heads("enc_slf_attn"), rows("enc_slf_fc"))
    e = _ffn(_sub(el, "pos_ffn."), e, rows("enc_ffn"), "enc_ffn")
    # Decoder.forward (:143-172)
    d = trg + w["decoder.position_enc.pos_table"][:, :n_t]
    if rows("dec_emb", True) is not None:                                          # Decoder.dropout (:152)
        d = d * rows("dec_emb", True)
    d = F.layer_norm(d, (512,), w["decoder.layer_norm.weight"], w["decoder.layer_norm.bias"], eps=1e-6)
    dl = _sub(w, "decoder.layer_stack.0.")
    d = _mha(_sub(dl, "slf_attn."), d, d, d, trg_mask, heads("dec_slf_attn", True), rows("dec_slf_fc", True))
    d = _mha(_sub(dl, "enc_attn."), d, e, e, src_mask, heads("dec_enc_attn"), rows("dec_enc_fc"))
    d = _ffn(_sub(dl, "pos_ffn."), d, rows("dec_ffn"), "dec_ffn")
    out = d.permute(0, 2, 1).contiguous().view(bp, 512, 8, 8)                      # :276-277
    out = F.conv2d(out, w["dec_trans.0.weight"], w["dec_trans.0.bias"])           # :278
    return (out, e) if return_enc else out


# ------------------------------------------------------------------------------------------------
# SKNet (lib/model/modules/blocks_coatt_transformer_sk.py:960-998) -- bug-compatible
# ------------------------------------------------------------------------------------------------
def sk_block(w, x):
    f1 = F.relu(F.conv2d(x, w["convs.0.0.weight"], w["convs.0.0.bias"], padding=0, groups=8))
    f3 = F.relu(F.conv2d(x, w["convs.1.0.weight"], w["convs.1.0.bias"], padding=1, groups=8))
    return f1 * f1 + f3 * f3          # v = f * f.expand_as(f); v.sum(dim=1)   (:980-983)


def sknet_forward(sd, x_props, x_query, dtype=torch.float32):
    w = _cast(sd, dtype)
    return sk_block(_sub(w, "sk_props."), x_props.to(dtype)), sk_block(_sub(w, "sk_query."), x_query.to(dtype))


# ------------------------------------------------------------------------------------------------
# RCNN_top = layer4, `_head_to_tail` (lib/model/faster_rcnn/resnet_coatt_transformer_sk.py:73-109,476-485)
# ------------------------------------------------------------------------------------------------
def _bn(w, p, x):
    return F.batch_norm(x, w[p + "running_mean"], w[p + "running_var"], w[p + "weight"], w[p + "bias"],
                        training=False, eps=1e-5)


def _bottleneck(w, x, stride, has_down, tag="l4"):
    out = _relu(_bn(w, "bn1.", F.conv2d(x, w["conv1.weight"], stride=stride)), tag + ".1")
    out = _relu(_bn(w, "bn2.", F.conv2d(out, w["conv2.weight"], padding=1)), tag + ".2")
    out = _bn(w, "bn3.", F.conv2d(out, w["conv3.weight"]))
    res = _bn(w, "downsample.1.", F.conv2d(x, w["downsample.0.weight"], stride=stride)) if has_down else x
    return _relu(out + res, tag + ".3")


def head_to_tail(sd, x, dtype=torch.float32):
    """sd: RCNN_top keys ('0.<block>.<...>').  x [G,1024,8,8] -> [G,2048]."""
    w = _cast(sd, dtype)
    x = x.to(dtype)
    call = _l4_calls[0]
    _l4_calls[0] += 1
    for i in range(3):
        x = _bottleneck(_sub(w, "0.%d." % i), x, 2 if i == 0 else 1, i == 0, "l4.%d.%d" % (call, i))
    return x.mean(3).mean(2)


# ------------------------------------------------------------------------------------------------
# whole head (lib/model/faster_rcnn/faster_rcnn_coatt_transformer_sk.py:273-337)
# ------------------------------------------------------------------------------------------------
def roi_align(feat, rois):
    """ROIAlign((7,7), 1/16, 0) via the C oracle (fp32, as the reference's CPU kernel)."""
    out = c_ops.roi_align_forward(feat.float().numpy(), rois.float().numpy(), 1.0 / 16.0, 7, 7, 0)
    return torch.from_numpy(out)


def head_forward(sd, non_img, non_qry, rois, dtype=torch.float32, roi_align_fn=None):
    """non_img [B,1024,H,W], non_qry [B,1024,8,8], rois [B,P,5] -> dict of outputs + intermediates.
    roi_align_fn: optional replacement for the C oracle (bench.py passes the reference's own CPU kernel)."""
    B, P = rois.shape[0], rois.shape[1]
    _l4_calls[0] = 0
    pooled = (roi_align_fn or roi_align)(non_img, rois.reshape(-1, 5))                 # :279
    ait, enc = ait_forward(_sub(sd, "transformer."), pooled, non_qry, dtype, True)     # :289
    sp, sq = sknet_forward(_sub(sd, "sk."), ait, non_qry, dtype)                       # :294
    top = _sub(sd, "RCNN_top.")
    pf = head_to_tail(top, sp, dtype)                                                  # :299
    qf = head_to_tail(top, sq, dtype)                                                  # :300
    w = _cast(sd, dtype)
    bbox = F.linear(pf, w["RCNN_bbox_pred.weight"], w["RCNN_bbox_pred.bias"])          # :318
    stack = torch.cat([pf.view(B, P, -1), qf.unsqueeze(1).repeat(1, P, 1)], dim=2).view(-1, 4096)   # :320-325
    score = F.linear(F.linear(stack, w["RCNN_cls_score.0.weight"], w["RCNN_cls_score.0.bias"]),
                     w["RCNN_cls_score.1.weight"], w["RCNN_cls_score.1.bias"])        # :335
    prob = F.softmax(score, 1)[:, 1]                                                   # :337
    return dict(pooled=pooled, enc_out=enc, ait_out=ait, sk_out=sp, feat=pf, qfeat=qf, score=score,
                cls_prob=prob.view(B, P, 1), bbox_pred=bbox.view(B, P, 4))


# ------------------------------------------------------------------------------------------------
# proposal tail (lib/model/rpn/proposal_layer.py:129-166), `>` NMS semantics of the CUDA path
# ------------------------------------------------------------------------------------------------
def propose_rois(proposals, scores, pre_nms_topN=6000, post_nms_topN=300, thr=0.7):
    import numpy as np
    B, n_total = scores.shape
    out = torch.zeros(B, post_nms_topN, 5)
    counts = []
    for i in range(B):
        s = scores[i].double().numpy()
        order = np.argsort(-s, kind="stable")
        if 0 < pre_nms_topN < scores.numel():
            order = order[:pre_nms_topN]
        boxes = proposals[i].float().numpy()[order]
        keep = c_ops.nms_sorted(boxes, thr, ge=False, max_keep=post_nms_topN)
        out[i, :, 0] = i
        out[i, : len(keep), 1:] = torch.from_numpy(boxes[keep])
        counts.append(len(keep))
    return out, counts


# ------------------------------------------------------------------------------------------------
# f1: the whole proposal layer (lib/model/rpn/proposal_layer.py:51-166, bbox_transform.py:77-133)
# ------------------------------------------------------------------------------------------------
def bbox_transform_inv(boxes, deltas):
    """bbox_transform.py:77-106 (class-agnostic: deltas [B, N, 4])."""
    widths = boxes[:, :, 2] - boxes[:, :, 0] + 1.0
    heights = boxes[:, :, 3] - boxes[:, :, 1] + 1.0
    ctr_x = boxes[:, :, 0] + 0.5 * widths
    ctr_y = boxes[:, :, 1] + 0.5 * heights
    pcx = deltas[:, :, 0] * widths + ctr_x
    pcy = deltas[:, :, 1] * heights + ctr_y
    pw = torch.exp(deltas[:, :, 2]) * widths
    ph = torch.exp(deltas[:, :, 3]) * heights
    return torch.stack([pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw, pcy + 0.5 * ph], dim=2)


def clip_boxes(boxes, im_info):
    """bbox_transform.py:125-133."""
    out = boxes.clone()
    for i in range(out.shape[0]):
        out[i, :, 0].clamp_(0, float(im_info[i, 1]) - 1)
        out[i, :, 1].clamp_(0, float(im_info[i, 0]) - 1)
        out[i, :, 2].clamp_(0, float(im_info[i, 1]) - 1)
        out[i, :, 3].clamp_(0, float(im_info[i, 0]) - 1)
    return out


def proposal_layer(rpn_cls_prob, rpn_bbox_pred, im_info, base_anchors, feat_stride=16, pre_nms_topN=6000,
                   post_nms_topN=300, nms_thresh=0.7, return_decoded=False):
    """_ProposalLayer.forward (proposal_layer.py:51-166).  NMS with the CUDA `>` rule (the path replaced)."""
    B, c2, H, W = rpn_cls_prob.shape
    A = c2 // 2
    scores = rpn_cls_prob[:, A:, :, :]
    sx = torch.arange(W, dtype=torch.float32) * feat_stride
    sy = torch.arange(H, dtype=torch.float32) * feat_stride
    yy, xx = torch.meshgrid(sy, sx, indexing="ij")
    shifts = torch.stack([xx.reshape(-1), yy.reshape(-1), xx.reshape(-1), yy.reshape(-1)], dim=1)
    K = shifts.shape[0]
    anchors = (base_anchors.view(1, A, 4) + shifts.view(K, 1, 4)).view(1, K * A, 4).expand(B, K * A, 4)
    deltas = rpn_bbox_pred.permute(0, 2, 3, 1).contiguous().view(B, -1, 4)
    scores = scores.permute(0, 2, 3, 1).contiguous().view(B, -1)
    proposals = clip_boxes(bbox_transform_inv(anchors, deltas), im_info)
    if return_decoded:
        return proposals, scores
    rois, _ = propose_rois(proposals, scores, pre_nms_topN, post_nms_topN, nms_thresh)
    return rois


# ------------------------------------------------------------------------------------------------
# f2: detection post-processing (test_net_voc.py:380-446), one unit at a time like the reference
# ------------------------------------------------------------------------------------------------
def detections(rois, cls_prob, bbox_pred, im_info, thresh=0.0, nms_thresh=0.3, max_per_image=100,
               stds=(0.1, 0.1, 0.2, 0.2), means=(0.0, 0.0, 0.0, 0.0), rescale=True):
    """-> list (per unit) of [n_det, 5] tensors (x1, y1, x2, y2, score), descending score."""
    import numpy as np
    out = []
    for b in range(rois.shape[0]):
        scores = cls_prob[b].reshape(-1)
        boxes = rois[b:b + 1, :, 1:5]
        deltas = bbox_pred[b].view(-1, 4) * torch.tensor(stds) + torch.tensor(means)          # :393-396
        pred = clip_boxes(bbox_transform_inv(boxes, deltas.view(1, -1, 4)), im_info[b:b + 1])  # :404-405
        if rescale:
            pred = pred / float(im_info[b, 2])                                                 # :410
        pred = pred[0]
        inds = torch.nonzero(scores > thresh).view(-1)                                         # :422
        if inds.numel() == 0:
            out.append(torch.zeros(0, 5))
            continue
        cls_scores, cls_boxes = scores[inds], pred[inds]
        cls_dets = torch.cat((cls_boxes, cls_scores.unsqueeze(1)), 1)
        _, order = torch.sort(cls_scores, stable=True, dim=0, descending=True)
        cls_dets = cls_dets[order]
        keep = c_ops.nms_sorted(cls_boxes[order].contiguous().numpy(), nms_thresh, False, 0)   # `>` rule
        dets = cls_dets[torch.from_numpy(np.asarray(keep).astype(np.int64))]
        if max_per_image > 0 and dets.shape[0] > max_per_image:                               # :437-446
            image_thresh = np.sort(dets[:, -1].numpy())[-max_per_image]
            dets = dets[dets[:, -1] >= float(image_thresh)]
        out.append(dets)
    return out


# ------------------------------------------------------------------------------------------------
# f3 (first half): the RPN head, inference path (lib/model/rpn/rpn.py:64-94)
# ------------------------------------------------------------------------------------------------
def rpn_forward(sd, base_feat, im_info, base_anchors, feat_stride=16, pre_nms_topN=6000, post_nms_topN=300,
                nms_thresh=0.7, dtype=torch.float32):
    """sd: RPN_Conv.*, RPN_cls_score.*, RPN_bbox_pred.* -> (rois, rpn_cls_prob, rpn_bbox_pred)."""
    w = _cast(sd, dtype)
    x = base_feat.to(dtype)
    conv1 = F.relu(F.conv2d(x, w["RPN_Conv.weight"], w["RPN_Conv.bias"], padding=1))          # :69
    score = F.conv2d(conv1, w["RPN_cls_score.weight"], w["RPN_cls_score.bias"])                 # :73
    B, c2, H, W = score.shape
    prob = F.softmax(score.contiguous().view(B, 2, c2 * H // 2, W), 1).view(B, c2, H, W)      # reshape(x, 2) :75-78
    bbox = F.conv2d(conv1, w["RPN_bbox_pred.weight"], w["RPN_bbox_pred.bias"])                  # :82
    rois = proposal_layer(prob.float(), bbox.float(), im_info, base_anchors, feat_stride, pre_nms_topN, post_nms_topN,
                          nms_thresh)
    return rois, prob, bbox


# ------------------------------------------------------------------------------------------------
# f3 (second half): CoAttention, 'division' normalisation (lib/model/modules/blocks_coatt_transformer_sk.py:60-112)
# ------------------------------------------------------------------------------------------------
def coattention_forward(sd, x_img, x_qry, dtype=torch.float32):
    """sd: emb.*, rho.*, phi.*, omega.0.*, omega.1.*, theta.0.*, theta.1.* -> (non_img, non_qry)."""
    w = _cast(sd, dtype)
    x_img, x_qry = x_img.to(dtype), x_qry.to(dtype)
    bz, _, h_i, w_i = x_img.shape
    _, _, h_q, w_q = x_qry.shape
    emb_img = F.conv2d(x_img, w["emb.weight"], w["emb.bias"]).view(bz, 512, -1).permute(0, 2, 1).contiguous()   # :70-71
    emb_qry = F.conv2d(x_qry, w["emb.weight"], w["emb.bias"]).view(bz, 512, -1).permute(0, 2, 1).contiguous()   # :73-74
    rho_qry = F.conv2d(x_qry, w["rho.weight"], w["rho.bias"]).view(bz, 512, -1).permute(0, 2, 1)                 # :76-77
    phi_img = F.conv2d(x_img, w["phi.weight"], w["phi.bias"]).view(bz, 512, -1)                                   # :79
    co = torch.matmul(rho_qry, phi_img)                                                                          # :81
    n_q, n_i = co.size(1), co.size(2)
    q2i = co / n_i                                                                                               # :91
    i2q = co.permute(0, 2, 1).contiguous() / n_q                                                                 # :92
    non_img = torch.matmul(i2q, emb_qry).permute(0, 2, 1).contiguous().view(bz, 512, h_i, w_i)                   # :99-101
    non_img = F.group_norm(F.conv2d(non_img, w["theta.0.weight"], w["theta.0.bias"]), 32, w["theta.1.weight"],
                           w["theta.1.bias"], eps=1e-5) + x_img                                                  # :102-104
    non_qry = torch.matmul(q2i, emb_img).permute(0, 2, 1).contiguous().view(bz, 512, h_q, w_q)                   # :106-108
    non_qry = F.group_norm(F.conv2d(non_qry, w["omega.0.weight"], w["omega.0.bias"]), 32, w["omega.1.weight"],
                           w["omega.1.bias"], eps=1e-5) + x_qry                                                  # :109-111
    return non_img, non_qry
