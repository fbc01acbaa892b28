"""Thin Python wrappers over the C ABI (one function per entry point).  Inputs are CUDA tensors;
outputs are allocated here with torch (caching allocator) and handed to the library as raw
pointers on torch's current stream.  No math happens on the Python side."""
import ctypes as C

import torch

from . import _lib as L


def _act_dtype(t):
    return L.dtype_enum(t.dtype)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("ait_b200: expected CUDA tensors (this path has no CPU implementation)")


# ---------------------------------------------------------------------------------------------
# layout
# ---------------------------------------------------------------------------------------------
def transpose_cs(src, to_channels_last, out_dtype=None, split_src=False, split_dst=False, round_tf32=False):
    """[G, C, S] -> [G, S, C] (to_channels_last) or [G, S, C] -> [G, C, S].
    split_dst: the (channels-last) destination is a two-plane bf16 matrix [G, S, hi C | lo C] (AITB_F32S);
    split_src: the (channels-last) source is one.  round_tf32: an fp32 destination is rounded to tf32 (nearest)."""
    lib = L.load()
    _need_cuda(src)
    src = src.contiguous()
    out_dtype = torch.bfloat16 if split_dst else (out_dtype or (torch.float32 if split_src else src.dtype))
    G = src.shape[0]
    if to_channels_last:
        Cc, S = src.shape[1], src.shape[2]
        dst = torch.empty((G, S, Cc * (2 if split_dst else 1)), device=src.device, dtype=out_dtype)
    else:
        S, Cc = src.shape[1], src.shape[2] // (2 if split_src else 1)
        dst = torch.empty((G, Cc, S), device=src.device, dtype=out_dtype)
    sdt = L.AITB_F32S if split_src else _act_dtype(src)
    ddt = L.AITB_F32S if split_dst else L.dtype_enum(out_dtype)
    step = 32768
    for g0 in range(0, G, step):
        gn = min(step, G - g0)
        L.check(lib.aitb_transpose_cs_round(L.ptr(src[g0:]), sdt, L.ptr(dst[g0:]), ddt, gn, Cc, S,
                                            1 if to_channels_last else 0, 1 if round_tf32 else 0, L.stream_ptr()))
    return dst


def split_planes(x, f16=False):
    """fp32 [..., C] -> two 16-bit planes [..., hi C | lo C] (host-side helper for tests / packing).  f16: the planes hold
    IEEE fp16 (the precision plan's operand format: 11-bit significand, saturating) inside the same bfloat16-typed container."""
    x = x.float()
    if f16:
        x = x.clamp(-65504.0, 65504.0)
        hi = x.to(torch.float16)
        lo = (x - hi.float()).to(torch.float16)
        return torch.cat([hi, lo], dim=-1).contiguous().view(torch.bfloat16)
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return torch.cat([hi, lo], dim=-1).contiguous()


def join_planes(x, f16=False):
    """[..., hi C | lo C] -> fp32 [..., C]  (f16: the container holds fp16 planes, see split_planes)."""
    if f16:
        x = x.contiguous().view(torch.float16)
    c = x.shape[-1] // 2
    return x[..., :c].float() + x[..., c:].float()


# ---------------------------------------------------------------------------------------------
# ROIAlign
# ---------------------------------------------------------------------------------------------
def roi_align_forward(feat_nhwc, rois, spatial_scale, pooled_h, pooled_w, sampling_ratio, token_major=False):
    """feat_nhwc [B, H, W, C] (fp32 | bf16), rois [K, 5] fp32 -> [K, C, ph, pw] or [K, ph*pw, C]."""
    lib = L.load()
    _need_cuda(feat_nhwc, rois)
    feat_nhwc = feat_nhwc.contiguous()
    rois = rois.contiguous().float()
    B, H, W, Cc = feat_nhwc.shape
    K = rois.shape[0]
    shape = (K, pooled_h * pooled_w, Cc) if token_major else (K, Cc, pooled_h, pooled_w)
    out = torch.empty(shape, device=feat_nhwc.device, dtype=feat_nhwc.dtype)
    step = 32768
    for k0 in range(0, K, step):
        kn = min(step, K - k0)
        L.check(lib.aitb_roi_align_forward(L.ptr(feat_nhwc), L.ptr(rois[k0:]), B, Cc, H, W, kn,
                                           float(spatial_scale), pooled_h, pooled_w, int(sampling_ratio),
                                           _act_dtype(feat_nhwc), 1 if token_major else 0, L.ptr(out[k0:]),
                                           L.stream_ptr()))
    return out


def roi_align_backward(grad, rois, spatial_scale, pooled_h, pooled_w, B, Cc, H, W, sampling_ratio):
    """grad [K, C, ph, pw] fp32 -> grad_input [B, C, H, W] fp32 (NCHW, like the reference)."""
    lib = L.load()
    _need_cuda(grad, rois)
    grad = grad.contiguous().float()
    rois = rois.contiguous().float()
    K = rois.shape[0]
    g_nhwc = torch.zeros((B, H, W, Cc), device=grad.device, dtype=torch.float32)
    step = 32768
    for k0 in range(0, K, step):
        kn = min(step, K - k0)
        L.check(lib.aitb_roi_align_backward(L.ptr(grad[k0:]), L.ptr(rois[k0:]), B, Cc, H, W, kn,
                                            float(spatial_scale), pooled_h, pooled_w, int(sampling_ratio),
                                            L.ptr(g_nhwc), L.stream_ptr()))
    return transpose_cs(g_nhwc.view(B, H * W, Cc), to_channels_last=False).view(B, Cc, H, W)


# ---------------------------------------------------------------------------------------------
# top-n selection + NMS
# ---------------------------------------------------------------------------------------------
_ws_cache = {}


def _workspace(nbytes, device, tag):
    key = (tag, device)
    buf = _ws_cache.get(key)
    # the usable part starts at the next 1024-byte boundary: keep that slack in the reuse test
    if buf is None or buf.numel() < int(nbytes) + 1024:
        buf = torch.empty(int(nbytes) + 1024, dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    off = (-buf.data_ptr()) % 1024
    return buf[off:]


def topk_desc(scores, n):
    """scores [B, n_total] fp32 -> order [B, n] int64 (descending score, ties by lower index)."""
    lib = L.load()
    _need_cuda(scores)
    scores = scores.contiguous().float()
    B, n_total = scores.shape
    n = min(n, n_total)
    order = torch.empty((B, n), dtype=torch.int64, device=scores.device)
    nbytes = lib.aitb_topk_workspace_bytes(B, n_total, n)
    ws = _workspace(nbytes, scores.device, "topk")
    L.check(lib.aitb_topk_desc(L.ptr(scores), B, n_total, n, L.ptr(order), L.ptr(ws), nbytes, L.stream_ptr()))
    return order


def nms_batched(boxes, order, thr, max_out, mode, want_rois=False):
    """boxes [B, n_total, 4] fp32, order [B, n] int64 or None -> (keep [B, max_out] int64, n_keep [B] int32, rois)."""
    lib = L.load()
    _need_cuda(boxes, order)
    boxes = boxes.contiguous().float()
    B, n_total = boxes.shape[0], boxes.shape[1]
    n = n_total if order is None else order.shape[1]
    if order is not None:
        order = order.contiguous()
    keep = torch.empty((B, max_out), dtype=torch.int64, device=boxes.device)
    n_keep = torch.empty((B,), dtype=torch.int32, device=boxes.device)
    rois = torch.empty((B, max_out, 5), dtype=torch.float32, device=boxes.device) if want_rois else None
    nbytes = lib.aitb_nms_workspace_bytes(B, n_total, n)
    ws = _workspace(nbytes, boxes.device, "nms")
    L.check(lib.aitb_nms_batched(L.ptr(boxes), L.ptr(order), B, n_total, n, float(thr), int(max_out), int(mode),
                                 L.ptr(keep), L.ptr(n_keep), L.ptr(rois), L.ptr(ws), nbytes, L.stream_ptr()))
    return keep, n_keep, rois


# ---------------------------------------------------------------------------------------------
# GEMM building block
# ---------------------------------------------------------------------------------------------
def _esize(dt):
    return 4 if dt == L.AITB_F32 else 2


def tiled_rows(H, W, B):
    """GEMM rows of the "tiled" map view: boxes of bx x by = 128 positions, ceil(H / by) boxes per image."""
    by = 128 // (64 if W <= 64 else 128)
    return B * ((H + by - 1) // by) * 128


def gemm(a, w, out, *, M, N, K, block_n, view="plain", lda=None, map_args=None, taps=1, group_c=0, flags=0,
         bias=None, res=None, ldr=0, res_div=1, res_rep=1, pos=None, pos_rows=1, gamma=None, beta=None,
         rows_in=None, rows_out=None, round_tf32=False, eps=1e-6, dual=False, bias2=None, split=False, ln_rstd=None,
         passes=0, in_f16=False, out_f16=False, res_f16=False, out_scale=0.0, ldo=None):
    """out = epilogue(A W^T).  `view`: "plain" (A is [M, lda]), "map" (A is a channels-last map,
    map_args = (C, S, s, stride, G): an s x s grid sampled with `stride` from an S x S map of C channels) or "tiled"
    (A is a channels-last map [B, H, W, C] of any size with W <= 128, map_args = (C, H, W, B), tiled by boxes of 128
    positions of one image -- the RPN 3x3 convolution; M must be tiled_rows(H, W, B), the output has B*H*W rows).
    split=True: AITB_F32S -- a / w / out / res are two-plane bf16 matrices (see split_planes); K, lda, ldr
    and the map's C are LOGICAL element counts."""
    lib = L.load()
    d = L.GemmDesc()
    dt = L.AITB_F32S if split else _act_dtype(a)
    eb = 4 if split else _esize(dt)            # bytes per logical element of a row
    cb = 2 if split else eb                    # bytes per TMA element
    d.dtype, d.M, d.N, d.k_per_tap, d.taps = dt, M, N, K, taps
    d.a.ptr = a.data_ptr()
    d.a_group_c = group_c
    if view == "plain":
        lda = lda or K
        d.a.dims[:] = [(lda if group_c else K * taps) + (lda if split else 0), M, 1, 1]
        d.a.strides[:] = [lda * eb, lda * eb * M, lda * eb * M]
        d.a.box[:] = [128 // cb, 128, 1, 1]
        d.a_m_dim, d.a_m_step = 1, 128
        d.a_lo_off = lda if split else 0
    elif view == "tiled":
        Cc, Hm, Wm, Bm = map_args
        bx = 64 if Wm <= 64 else 128
        if Wm > 128 or M != tiled_rows(Hm, Wm, Bm):
            raise RuntimeError("ait_b200.gemm: tiled map view needs W <= 128 and M = tiled_rows(H, W, B)")
        d.a.dims[:] = [Cc * (2 if split else 1), Wm, Hm, Bm]
        d.a.strides[:] = [Cc * eb, Wm * Cc * eb, Hm * Wm * Cc * eb]
        d.a.box[:] = [128 // cb, bx, 128 // bx, 1]
        d.a_m_dim, d.a_m_step = 2, 1
        d.a_lo_off = Cc if split else 0
        d.map_w, d.map_h = Wm, Hm
    else:
        Cc, S, s, stride, G = map_args
        d.a.dims[:] = [Cc * (2 if split else 1), s, s, G]
        d.a.strides[:] = [stride * Cc * eb, stride * S * Cc * eb, S * S * Cc * eb]
        d.a.box[:] = [128 // cb, s, s, 128 // (s * s)]
        d.a_m_dim, d.a_m_step = 3, 128 // (s * s)
        d.a_lo_off = Cc if split else 0
    if taps == 9:
        for ky in range(3):
            for kx in range(3):
                d.tap_dx[ky * 3 + kx] = kx - 1
                d.tap_dy[ky * 3 + kx] = ky - 1
    d.w = w.data_ptr()
    d.block_n = block_n
    d.flags = flags
    d.out = out.data_ptr()
    d.ldo = ldo or out.shape[-1] // (2 if split else 1)     # ldo: `out` is a column block of a wider row-major buffer
    d.rows_in = rows_in or M
    d.rows_out = rows_out or M
    d.bias = 0 if bias is None else bias.data_ptr()
    d.res = 0 if res is None else res.data_ptr()
    d.ldr, d.res_div, d.res_rep = ldr, res_div, res_rep
    d.pos = 0 if pos is None else pos.data_ptr()
    d.pos_rows = pos_rows
    d.gamma = 0 if gamma is None else gamma.data_ptr()
    d.beta = 0 if beta is None else beta.data_ptr()
    d.eps = eps
    d.round_tf32 = 1 if round_tf32 else 0
    d.dual = 1 if dual else 0
    d.bias2 = 0 if bias2 is None else bias2.data_ptr()
    d.ln_rstd = 0 if ln_rstd is None else ln_rstd.data_ptr()
    d.passes, d.in_f16, d.out_f16, d.res_f16 = int(passes), int(bool(in_f16)), int(bool(out_f16)), int(bool(res_f16))
    d.out_scale = float(out_scale)                # the accumulator is multiplied by it before bias / activation (0 = unset = 1)
    L.check(lib.aitb_gemm(C.byref(d), L.stream_ptr()))
    return out


def fc_ln(a, w_fc, res, gamma, beta, out, *, M, split=False, rows_in=None, rows_out=None, res_row_m=False, res_div=1,
          res_rep=1, eps=1e-6):
    """out = LayerNorm(a w_fc^T + res) (N = 512, K = 64) by the streaming kernel; bf16 storage or split planes."""
    lib = L.load()
    _need_cuda(a, w_fc, res, gamma, beta, out)
    L.check(lib.aitb_fc_ln(L.AITB_F32S if split else L.AITB_BF16, L.ptr(a), L.ptr(w_fc), L.ptr(res), L.ptr(gamma), L.ptr(beta),
                           float(eps), L.ptr(out), M, rows_in or M, rows_out or M, 1 if res_row_m else 0, res_div, res_rep,
                           L.stream_ptr()))
    return out


def attn_core(q, ldq, q_rep, k, v, ldkv, w_sk, b_sk, G, mask_mode, n_keys, out, split=False):
    """split=True: q / k / v / out are two-plane bf16 matrices, ldq / ldkv their LOGICAL row widths."""
    lib = L.load()
    L.check(lib.aitb_attn_core(L.ptr(q), ldq, q_rep, L.ptr(k), L.ptr(v), ldkv, L.ptr(w_sk), L.ptr(b_sk), G,
                               mask_mode, n_keys, L.AITB_F32S if split else _act_dtype(q), L.ptr(out),
                               L.stream_ptr()))
    return out


def pool_heads(top, P, qfeat=None, w_bbox=None, b_bbox=None, w1=None, b1=None, w2=None, b2=None, want_feat=True,
               split=False):
    """top [G, 16, 2048] -> (feat [G, 2048] | None, bbox [G, 4] | None, cls_prob [G] | None)."""
    lib = L.load()
    G = top.shape[0]
    dev = top.device
    feat = torch.empty((G, 2048), dtype=torch.float32, device=dev) if want_feat else None
    heads = w_bbox is not None
    bbox = torch.empty((G, 4), dtype=torch.float32, device=dev) if heads else None
    cls = torch.empty((G,), dtype=torch.float32, device=dev) if heads else None
    L.check(lib.aitb_pool_heads(L.ptr(top), L.AITB_F32S if split else _act_dtype(top), G, P, L.ptr(qfeat), L.ptr(w_bbox), L.ptr(b_bbox),
                                L.ptr(w1), L.ptr(b1), L.ptr(w2), L.ptr(b2), L.ptr(feat), L.ptr(bbox), L.ptr(cls),
                                L.stream_ptr()))
    return feat, bbox, cls


# ---------------------------------------------------------------------------------------------
# training building blocks (fp32 storage, tf32 tensor-core math)
# ---------------------------------------------------------------------------------------------
def wgrad(dy, x, dw=None, N=None, K=None):
    """dw[N, K] += dy[:, :N]^T x[:, :K] for row-major dy [M, ldy], x [M, ldx], both fp32 (tf32 MMAs) or both bf16; dw
    fp32 (zero-initialised when not given)."""
    lib = L.load()
    _need_cuda(dy, x)
    if dy.dtype != x.dtype or dy.dtype not in (torch.float32, torch.bfloat16) or dy.stride(-1) != 1 or x.stride(-1) != 1:
        raise RuntimeError("ait_b200.wgrad: row-major operands of one dtype (float32 or bfloat16) required")
    M = dy.shape[0]
    N = N or dy.shape[1]
    K = K or x.shape[1]
    if dw is None:
        dw = torch.zeros((N, K), dtype=torch.float32, device=dy.device)
    fn = lib.aitb_wgrad if dy.dtype == torch.float32 else lib.aitb_wgrad_bf16
    L.check(fn(L.ptr(dy), dy.stride(0), L.ptr(x), x.stride(0), M, N, K, L.ptr(dw), dw.stride(0), L.stream_ptr()))
    return dw


def wgrad_conv(dy, x_map, G, S, Cc, N, groups=1, taps=9, dw=None):
    """Weight gradient of a (grouped) 1x1 / 3x3 "same" convolution without im2col: dy [G*S*S, >=N] row-major fp32,
    x_map the channels-last input [G*S*S, Cc]; -> dw [N, taps * Cc / groups] (tap-major), accumulated when given."""
    lib = L.load()
    _need_cuda(dy, x_map)
    if dy.dtype != torch.float32 or x_map.dtype != torch.float32 or dy.stride(-1) != 1 or not x_map.is_contiguous():
        raise RuntimeError("ait_b200.wgrad_conv: float32 row-major dy and a contiguous channels-last map required")
    if dy.shape[0] != G * S * S or x_map.numel() != G * S * S * Cc:
        raise RuntimeError("ait_b200.wgrad_conv: dy / x_map do not match G=%d S=%d C=%d" % (G, S, Cc))
    K = taps * (Cc // groups)
    if dw is None:
        dw = torch.zeros((N, K), dtype=torch.float32, device=dy.device)
    L.check(lib.aitb_wgrad_conv(L.ptr(dy), dy.stride(0), L.ptr(x_map), G, S, Cc, N, groups, taps, L.ptr(dw), dw.stride(0),
                                L.stream_ptr()))
    return dw


def ln_bwd(g, y, gamma, beta, rstd, grp=64, valid=64):
    """LayerNorm backward from the saved output y [rows, 512] and 1/sigma [rows] -> (dx [rows/grp*valid, 512], dgamma, dbeta)."""
    lib = L.load()
    rows = g.shape[0]
    dx = torch.empty((rows // grp * valid, 512), dtype=torch.float32, device=g.device)
    dgamma = torch.zeros(512, dtype=torch.float32, device=g.device)
    dbeta = torch.zeros(512, dtype=torch.float32, device=g.device)
    L.check(lib.aitb_ln_bwd(L.ptr(g), L.ptr(y), L.ptr(gamma), L.ptr(beta), L.ptr(rstd), rows, grp, valid, L.ptr(dx),
                            L.ptr(dgamma), L.ptr(dbeta), L.stream_ptr()))
    return dx, dgamma, dbeta


def colsum(x, cols=None):
    lib = L.load()
    cols = cols or x.shape[1]
    out = torch.zeros(cols, dtype=torch.float32, device=x.device)
    L.check(lib.aitb_colsum(L.ptr(x), x.stride(0), x.shape[0], cols, L.ptr(out), L.stream_ptr()))
    return out


def bsum(x):
    """[B, P, L] -> [B, L]"""
    lib = L.load()
    B, P, Ln = x.shape
    out = torch.empty((B, Ln), dtype=torch.float32, device=x.device)
    L.check(lib.aitb_bsum(L.ptr(x.contiguous()), B, P, Ln, L.ptr(out), L.stream_ptr()))
    return out


def attn_bwd(q, ldq, q_rep, k, v, ldkv, w_sk, b_sk, dout, G, mask_mode, n_keys):
    """-> (dq [G*64, 512] per pair, dk [G*64, 512], dv [G*64, 512], dz [G, 512], s [G, 64])"""
    lib = L.load()
    dev = dout.device
    dq = torch.empty((G * 64, 512), dtype=torch.float32, device=dev)
    dk = torch.empty((G * 64, 512), dtype=torch.float32, device=dev)
    dv = torch.empty((G * 64, 512), dtype=torch.float32, device=dev)
    dz = torch.empty((G, 512), dtype=torch.float32, device=dev)
    s = torch.empty((G, 64), dtype=torch.float32, device=dev)
    L.check(lib.aitb_attn_bwd(L.ptr(q), ldq, q_rep, L.ptr(k), L.ptr(v), ldkv, L.ptr(w_sk), L.ptr(b_sk), L.ptr(dout), G,
                              mask_mode, n_keys, L.ptr(dq), 512, L.ptr(dk), L.ptr(dv), 512, L.ptr(dz), L.ptr(s),
                              L.stream_ptr()))
    return dq, dk, dv, dz, s


def launch_count(reset=False):
    return int(L.load(check_device=False).aitb_launch_count(1 if reset else 0))


# every wrapper above runs on the device of its first CUDA tensor argument (one device guard, in one place)
L.guard_module_functions(globals(), __name__, skip=("split_planes", "join_planes", "launch_count"))
