"""The RPN head (`_RPN.forward`, lib/model/rpn/rpn.py:66-83: RPN_Conv 3x3 + ReLU, RPN_cls_score and RPN_bbox_pred 1x1) as a
differentiable function on the device -- row f3's training path.  The reference obtains the backward from torch autograd;
here forward and backward are composed from the library's building blocks (fp32 storage, tf32 tensor-core math):

  forward   the 3x3 convolution as ONE tcgen05 GEMM over the channels-last C4 map of any H x W (boxes of 128 positions, nine
            shifted zero-filled TMA boxes per K chunk: no im2col), bias + ReLU in the epilogue; both 1x1 heads as one GEMM
            against the stacked, zero-padded [128, 512] weight
  backward  heads: dgrad with the ReLU mask fused (`AITB_EPI_RELU_MASK`), wgrad `aitb_wgrad`, biases `aitb_colsum`;
            3x3 input gradient: the same conv GEMM with the nine taps flipped and transposed; 3x3 weight gradient: nine
            `aitb_wgrad` launches on row-shifted views of zero-bordered copies of the gradient and input maps (`aitb_map_pad`)
            -- in the padded layout a tap is a constant row offset, and the border rows contribute zeros

Returns the raw maps the reference's losses consume (rpn_cls_score [B,2A,H,W], rpn_bbox_pred [B,4A,H,W]); gradients flow to
base_feat and the six parameters.  No CPU / eager fallback.
"""
import torch

from . import _lib as L
from . import ops
from .packing import round_to_tf32
from .top_train import dgrad_weight_3x3

NPAD = 128       # head GEMM width: 2A + 4A <= 128 columns, zero-padded (also the wgrad's 128-row tile)
_last_conv1_for_tests = None     # the last forward's post-ReLU conv map [B*H*W, 512] (the parity test takes the ReLU decisions from it)


def _pad_map(x, B, H, W, Cc, guard):
    """[B*H*W, C] channels-last -> zero-bordered [B, H+2, W+2, C] flattened, with `guard` zero rows before and after."""
    lib = L.load()
    Mp = B * (H + 2) * (W + 2)
    buf = torch.zeros((guard + Mp + guard, Cc), dtype=torch.float32, device=x.device)
    L.check(lib.aitb_map_pad(L.ptr(x), B, H, W, Cc, L.ptr(buf[guard:]), L.stream_ptr()))
    return buf, Mp


class _RPNHeadFn(torch.autograd.Function):
    """forward(base_feat [B,1024,H,W], conv_w [512,1024,3,3], conv_b, cls_w [2A,512,1,1], cls_b, box_w [4A,512,1,1], box_b)
    -> (rpn_cls_score [B,2A,H,W], rpn_bbox_pred [B,4A,H,W])."""

    @staticmethod
    @L.on_tensor_device
    def forward(ctx, base_feat, conv_w, conv_b, cls_w, cls_b, box_w, box_b):
        ops._need_cuda(base_feat, conv_w, conv_b, cls_w, cls_b, box_w, box_b)
        B, c, H, W = base_feat.shape
        A2, A4 = cls_w.shape[0], box_w.shape[0]
        if c != 1024 or tuple(conv_w.shape) != (512, 1024, 3, 3) or A2 + A4 > NPAD or W > 128:
            raise RuntimeError("rpn_head_train: expected base_feat [B,1024,H,W<=128], RPN_Conv [512,1024,3,3], <= 21 anchors")
        dev = base_feat.device
        rows = B * H * W
        wc = round_to_tf32(conv_w.detach().float().permute(0, 2, 3, 1).reshape(512, 9 * 1024).contiguous())   # tap-major
        wh = torch.zeros((NPAD, 512), dtype=torch.float32, device=dev)
        wh[:A2] = cls_w.detach().float().flatten(1)
        wh[A2:A2 + A4] = box_w.detach().float().flatten(1)
        wh = round_to_tf32(wh)
        bh = torch.zeros(NPAD, dtype=torch.float32, device=dev)
        bh[:A2] = cls_b.detach().float()
        bh[A2:A2 + A4] = box_b.detach().float()
        X = ops.transpose_cs(base_feat.detach().float().contiguous().view(B, 1024, H * W), True, round_tf32=True).view(rows, 1024)
        conv1 = torch.empty((rows, 512), dtype=torch.float32, device=dev)
        ops.gemm(X, wc, conv1, M=ops.tiled_rows(H, W, B), N=512, K=1024, block_n=256, view="tiled", map_args=(1024, H, W, B),
                 taps=9, flags=L.EPI_BIAS | L.EPI_RELU, bias=conv_b.detach().float().contiguous(), round_tf32=True)
        heads = torch.empty((rows, NPAD), dtype=torch.float32, device=dev)
        ops.gemm(conv1, wh, heads, M=rows, N=NPAD, K=512, block_n=128, flags=L.EPI_BIAS, bias=bh)
        h3 = heads.view(B, H * W, NPAD)
        score = ops.transpose_cs(h3[:, :, :A2].contiguous(), False).view(B, A2, H, W)
        bbox = ops.transpose_cs(h3[:, :, A2:A2 + A4].contiguous(), False).view(B, A4, H, W)
        ctx.dims = (B, H, W, A2, A4)
        ctx.keep = (X, conv1, wc, wh)
        global _last_conv1_for_tests
        _last_conv1_for_tests = conv1
        ctx.in_dtype = base_feat.dtype
        return score, bbox

    @staticmethod
    @L.on_tensor_device
    def backward(ctx, g_score, g_bbox):
        if ctx.keep is None:
            raise RuntimeError("ait_b200._RPN: backward a second time: the saved activations were freed after the first "
                               "backward (retain_graph=True is not supported; run the forward again)")
        X, conv1, wc, wh = ctx.keep
        B, H, W, A2, A4 = ctx.dims
        dev = X.device
        rows = B * H * W
        # gradient of the stacked head output [rows, 128] (channels-last; the padding columns stay zero)
        gH = torch.zeros((B, H * W, NPAD), dtype=torch.float32, device=dev)
        gH[:, :, :A2] = ops.transpose_cs(g_score.contiguous().float().view(B, A2, H * W), True, round_tf32=True)
        gH[:, :, A2:A2 + A4] = ops.transpose_cs(g_bbox.contiguous().float().view(B, A4, H * W), True, round_tf32=True)
        gH = gH.view(rows, NPAD)
        dWh = ops.wgrad(gH, conv1, N=NPAD, K=512)
        dbh = ops.colsum(gH)
        gC = torch.empty((rows, 512), dtype=torch.float32, device=dev)          # through the heads and the ReLU
        ops.gemm(gH, wh.t().contiguous(), gC, M=rows, N=512, K=NPAD, block_n=256, flags=L.EPI_RELU_MASK, res=conv1, ldr=512,
                 round_tf32=True)
        d_conv_b = ops.colsum(gC)
        # 3x3 weight gradient: dW_tap[n, c] = sum over positions of gC[b, y, x, n] * X[b, y + ky - 1, x + kx - 1, c]
        guard = W + 3
        gCp, Mp = _pad_map(gC, B, H, W, 512, 0)
        Xp, _ = _pad_map(X, B, H, W, 1024, guard)
        dWc = torch.zeros((9, 512, 1024), dtype=torch.float32, device=dev)
        for ky in range(3):
            for kx in range(3):
                delta = (ky - 1) * (W + 2) + (kx - 1)
                ops.wgrad(gCp, Xp[guard + delta:guard + delta + Mp], dw=dWc[ky * 3 + kx], N=512, K=1024)
        # 3x3 input gradient: the convolution of the gradient map with the flipped, transposed taps
        wcd = dgrad_weight_3x3(wc, 512, 1024)                                    # [1024, 9 * 512]
        gX = torch.empty((rows, 1024), dtype=torch.float32, device=dev)
        ops.gemm(gC, wcd, gX, M=ops.tiled_rows(H, W, B), N=1024, K=512, block_n=256, view="tiled", map_args=(512, H, W, B),
                 taps=9)
        d_feat = ops.transpose_cs(gX.view(B, H * W, 1024), False).view(B, 1024, H, W)
        ctx.keep = None
        d_conv_w = dWc.view(3, 3, 512, 1024).permute(2, 3, 0, 1).contiguous()   # [512, 1024, ky, kx]
        return (d_feat.to(ctx.in_dtype), d_conv_w, d_conv_b, dWh[:A2].reshape(A2, 512, 1, 1).clone(), dbh[:A2].clone(),
                dWh[A2:A2 + A4].reshape(A4, 512, 1, 1).clone(), dbh[A2:A2 + A4].clone())


def rpn_head_train(rpn, base_feat):
    """Differentiable RPN head over the module's own Parameters (the reference's `_RPN` or ait_b200.rpn._RPN: same names)
    -> (rpn_cls_score [B,2A,H,W], rpn_bbox_pred [B,4A,H,W])."""
    return _RPNHeadFn.apply(base_feat, rpn.RPN_Conv.weight, rpn.RPN_Conv.bias, rpn.RPN_cls_score.weight,
                            rpn.RPN_cls_score.bias, rpn.RPN_bbox_pred.weight, rpn.RPN_bbox_pred.bias)
