"""`_head_to_tail` (ResNet-50 layer4 with frozen BatchNorm + the 4x4 spatial mean) as a differentiable function on the
device: lib/model/faster_rcnn/resnet_coatt_transformer_sk.py:73-109 (Bottleneck), :416 (RCNN_top), :429-435 (frozen
BN), :476-485 (_head_to_tail).  The reference obtains the backward from torch autograd; here forward and backward are
composed from the library's building blocks (fp32 storage, tf32 tensor-core math, like the AIT training step):

  forward   the conv GEMMs of `HeadEngine.top_forward` (strided / nine-tap TMA views, BN folded into weight and bias,
            bias + ReLU / residual + ReLU epilogues), every activation the backward needs kept
  dgrad     the same tcgen05 GEMM with a transposed weight: 1x1 -> W^T; 3x3 -> the nine taps flipped and transposed
            (a 3x3 convolution of the gradient map); ReLU masks and the residual sum fused into the epilogues
  wgrad     `aitb_wgrad` (dW += dY^T X on MN-major operands); the 3x3 weight gradient is ONE `aitb_wgrad_conv` launch reading
            the saved 4x4 map through nine shifted, zero-filled TMA boxes -- tap-major, the layout the weights are packed in
  stride 2  the first bottleneck's 1x1 convolutions read every second position: gather / zero-fill scatter kernels

Gradients are returned for the ten convolution weights (BatchNorm is frozen: `set_bn_fix`, its parameters do not
train) and for the input map.  No CPU / eager fallback.
"""
import ctypes as C

import torch

from . import _lib as L
from . import ops
from .packing import round_to_tf32

_F = L.EPI_BIAS | L.EPI_RELU
_last_saved_for_tests = None      # the last call's saved activations
_saved_log_for_tests = None       # tests set this to a list: every call appends its saved activations (props call, then query call)


def _tap_major(w):
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)


def _call(fn, *args):
    L.check(fn(*args, L.stream_ptr()))


def _relu_bwd(dy, y):
    out = torch.empty_like(dy)
    _call(L.load().aitb_relu_bwd, L.ptr(dy), L.ptr(y), L.ptr(out), C.c_size_t(dy.numel()))
    return out


def _subsample(x, G, S, s, stride, Cc):
    out = torch.empty((G * s * s, Cc), dtype=torch.float32, device=x.device)
    _call(L.load().aitb_map_subsample, L.ptr(x), G, S, s, stride, Cc, L.ptr(out))
    return out


def _upsample(x, G, S, s, stride, Cc):
    out = torch.empty((G * S * S, Cc), dtype=torch.float32, device=x.device)
    _call(L.load().aitb_map_upsample, L.ptr(x), G, S, s, stride, Cc, L.ptr(out))
    return out


def dgrad_weight_3x3(w_packed, c_out, c_in):
    """Tap-major packed 3x3 weight [c_out, 9*c_in] -> the packed weight [c_in, 9*c_out] of the input-gradient convolution
    (taps flipped, in / out swapped).  Pure tensor re-indexing (host logic; CPU-testable)."""
    return w_packed.view(c_out, 9, c_in).flip(1).permute(2, 1, 0).contiguous().view(c_in, 9 * c_out)


class _HeadToTailFn(torch.autograd.Function):
    """forward(x_nchw [G,1024,8,8] (cl_in: the channels-last, tf32-rounded [G,64,1024] map `sk_train` hands over), consts,
    cl_in, *conv_weights) -> feat [G,2048].
    consts: per block a dict of BN scales / folded biases (not differentiated); conv_weights in the order
    (b0.conv1, b0.conv2, b0.conv3, b0.downsample.0, b1.conv1, b1.conv2, b1.conv3, b2.conv1, b2.conv2, b2.conv3)."""

    @staticmethod
    @L.on_tensor_device
    def forward(ctx, x_nchw, consts, cl_in, *weights):
        ops._need_cuda(x_nchw, *weights)
        G = x_nchw.shape[0]
        if tuple(x_nchw.shape[1:]) != ((64, 1024) if cl_in else (1024, 8, 8)) or len(weights) != 10:
            raise RuntimeError("head_to_tail_train: expected x [G,1024,8,8] and the ten layer4 convolution weights")
        dev = x_nchw.device
        M = G * 16
        if cl_in:
            x0 = x_nchw.detach().contiguous().float().view(G * 64, 1024)
        else:
            x0 = ops.transpose_cs(x_nchw.detach().contiguous().float().reshape(G, 1024, 64), True, out_dtype=torch.float32,
                                  round_tf32=True).view(G * 64, 1024)
        names = [("conv1", "conv2", "conv3", "down"), ("conv1", "conv2", "conv3"), ("conv1", "conv2", "conv3")]
        wi = iter(weights)
        packed, saved = [], []
        cur = x0
        for b in range(3):
            wd = {}
            for n in names[b]:
                w = next(wi).detach().float()
                wd[n] = round_to_tf32((_tap_major(w) * consts[b][n + "_scale"].view(-1, 1)).contiguous())   # BN folded
            packed.append(wd)
            cin = 1024 if b == 0 else 2048
            o1 = torch.empty((M, 512), dtype=torch.float32, device=dev)
            o2 = torch.empty((M, 512), dtype=torch.float32, device=dev)
            y = torch.empty((M, 2048), dtype=torch.float32, device=dev)
            if b == 0:
                ops.gemm(cur, wd["conv1"], o1, M=M, N=512, K=1024, block_n=256, view="map", map_args=(1024, 8, 4, 2, G),
                         flags=_F, bias=consts[b]["conv1_bias"], round_tf32=True)
                res = torch.empty((M, 2048), dtype=torch.float32, device=dev)
                ops.gemm(cur, wd["down"], res, M=M, N=2048, K=1024, block_n=256, view="map", map_args=(1024, 8, 4, 2, G),
                         flags=L.EPI_BIAS, bias=consts[b]["down_bias"], round_tf32=True)
            else:
                ops.gemm(cur, wd["conv1"], o1, M=M, N=512, K=cin, block_n=256, flags=_F, bias=consts[b]["conv1_bias"],
                         round_tf32=True)
                res = cur
            ops.gemm(o1, wd["conv2"], o2, M=M, N=512, K=512, block_n=256, view="map", map_args=(512, 4, 4, 1, G), taps=9,
                     flags=_F, bias=consts[b]["conv2_bias"], round_tf32=True)
            ops.gemm(o2, wd["conv3"], y, M=M, N=2048, K=512, block_n=256, flags=L.EPI_BIAS | L.EPI_RES | L.EPI_RES_RELU,
                     bias=consts[b]["conv3_bias"], res=res, ldr=2048, round_tf32=True)
            saved.append((o1, o2, y))
            cur = y
        feat, _, _ = ops.pool_heads(cur.view(G, 16, 2048), 1)
        ctx.G, ctx.consts, ctx.packed, ctx.saved, ctx.x0, ctx.cl_in = G, consts, packed, saved, x0, cl_in
        global _last_saved_for_tests
        _last_saved_for_tests = saved   # the parity test rebuilds the reference with exactly these ReLU masks
        if _saved_log_for_tests is not None:
            _saved_log_for_tests.append(saved)
        ctx.wshapes = [tuple(w.shape) for w in weights]
        return feat

    @staticmethod
    @L.on_tensor_device
    def backward(ctx, d_feat):
        lib = L.load()
        G, consts, packed, saved, x0 = ctx.G, ctx.consts, ctx.packed, ctx.saved, ctx.x0
        dev = x0.device
        M = G * 16
        d_feat = d_feat.contiguous().float()
        dy = torch.empty((G, 16, 2048), dtype=torch.float32, device=dev)
        _call(lib.aitb_mean_pool_backward, L.ptr(d_feat), G, L.ptr(dy))
        dy = dy.view(M, 2048)
        grads = {}

        def wg(dyv, xv, N, K):      # both operands are already tf32-rounded where they were produced
            return ops.wgrad(dyv, xv, N=N, K=K)

        for b in (2, 1, 0):
            o1, o2, y = saved[b]
            wd = packed[b]
            cin = 1024 if b == 0 else 2048
            xin = _subsample(x0, G, 8, 4, 2, 1024) if b == 0 else saved[b - 1][2]
            g = _relu_bwd(dy, y)                                             # through the block's final ReLU (tf32-rounded)
            # conv3 (1x1, 512 -> 2048)
            grads[(b, "conv3")] = wg(g, o2, 2048, 512)
            d_o2 = torch.empty((M, 512), dtype=torch.float32, device=dev)
            ops.gemm(g, wd["conv3"].t().contiguous(), d_o2, M=M, N=512, K=2048, block_n=256,
                     flags=L.EPI_RELU_MASK, res=o2, ldr=512, round_tf32=True)
            # conv2 (3x3, 512 -> 512): weight gradient straight from the saved 4x4 map (nine shifted TMA boxes, no im2col)
            grads[(b, "conv2")] = ops.wgrad_conv(d_o2, o1, G, 4, 512, 512, groups=1, taps=9)
            w2d = dgrad_weight_3x3(wd["conv2"], 512, 512)
            d_o1 = torch.empty((M, 512), dtype=torch.float32, device=dev)
            ops.gemm(d_o2, w2d, d_o1, M=M, N=512, K=512, block_n=256, view="map", map_args=(512, 4, 4, 1, G), taps=9,
                     flags=L.EPI_RELU_MASK, res=o1, ldr=512, round_tf32=True)
            # conv1 (1x1, cin -> 512) and the shortcut
            grads[(b, "conv1")] = wg(d_o1, xin, 512, cin)
            dx = torch.empty((M, cin), dtype=torch.float32, device=dev)
            w1t = wd["conv1"].t().contiguous()
            if b > 0:   # identity shortcut: dx = d_o1 W1 + g
                ops.gemm(d_o1, w1t, dx, M=M, N=cin, K=512, block_n=256, flags=L.EPI_RES, res=g, ldr=2048, round_tf32=True)
                dy = dx
            else:       # projection shortcut (1x1, stride 2) on the same strided rows
                grads[(b, "down")] = wg(g, xin, 2048, 1024)
                ops.gemm(d_o1, w1t, dx, M=M, N=1024, K=512, block_n=256, round_tf32=False)
                ops.gemm(g, wd["down"].t().contiguous(), dx, M=M, N=1024, K=2048, block_n=256,
                         flags=L.EPI_ACCUM, round_tf32=False)
                dx0 = _upsample(dx, G, 8, 4, 2, 1024)                        # [G*64, 1024], zeros off the stride-2 grid
        if ctx.cl_in:
            dx_nchw = dx0.view(G, 64, 1024)
        else:
            dx_nchw = ops.transpose_cs(dx0.view(G, 64, 1024), False, out_dtype=torch.float32).view(G, 1024, 8, 8)
        order = [(0, "conv1"), (0, "conv2"), (0, "conv3"), (0, "down"), (1, "conv1"), (1, "conv2"), (1, "conv3"),
                 (2, "conv1"), (2, "conv2"), (2, "conv3")]
        outs = []
        for (b, n), shp in zip(order, ctx.wshapes):
            gw = grads[(b, n)] * consts[b][n + "_scale"].view(-1, 1)         # d(folded) -> d(conv.weight): BN scale per row
            o, i, kh, kw = shp
            outs.append(gw.view(o, kh, kw, i).permute(0, 3, 1, 2).contiguous())
        return (dx_nchw, None, None) + tuple(outs)


def head_to_tail_train(RCNN_top, pool5, channels_last=False):
    """Differentiable `_head_to_tail(pool5)` for the reference's `RCNN_top = nn.Sequential(resnet.layer4)` with frozen
    BatchNorm: pool5 [G,1024,8,8] -> [G,2048].  Gradients flow to pool5 and to the ten convolution weights.
    channels_last: pool5 is the [G,64,1024] token-major, tf32-rounded map of `sk_train.sknet_train(..., channels_last_out=True)`."""
    layer4 = RCNN_top[0]
    consts, weights = [], []
    for blk in layer4:
        c = {}
        convs = [("conv1", blk.conv1, blk.bn1), ("conv2", blk.conv2, blk.bn2), ("conv3", blk.conv3, blk.bn3)]
        if blk.downsample is not None:
            convs.append(("down", blk.downsample[0], blk.downsample[1]))
        order = ["conv1", "conv2", "conv3"] + (["down"] if blk.downsample is not None else [])
        byname = {n: (cv, bn) for n, cv, bn in convs}
        for n in order:
            cv, bn = byname[n]
            scale = (bn.weight.detach().float() / torch.sqrt(bn.running_var.float() + bn.eps)).contiguous()
            c[n + "_scale"] = scale
            c[n + "_bias"] = (bn.bias.detach().float() - bn.running_mean.float() * scale).contiguous()
            weights.append(cv.weight)
        consts.append(c)
    return _HeadToTailFn.apply(pool5, consts, bool(channels_last), *weights)
