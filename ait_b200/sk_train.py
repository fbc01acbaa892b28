"""`SKNet` / `SKBlock` (lib/model/modules/blocks_coatt_transformer_sk.py:915-998) as a differentiable function on the
device.  What the reference's forward returns is `relu(conv1x1_g8(x))**2 + relu(conv3x3_g8(x))**2` (:973-984: the
selective-kernel attention `a` is computed and then discarded, `v = f * f`), so `fc` / `sk` receive no gradient
(autograd gives them None); the two grouped convolutions (groups = 8, bias) train.  The reference obtains the backward
from torch autograd; here forward and backward are composed from the library's building blocks (fp32 storage, tf32
tensor-core math, like the AIT and layer-4 training steps):

  forward   two grouped tcgen05 GEMMs on the channels-last 8x8 map (group = n-tile of 128 output channels reading its
            own 128 input channels; the 3x3 branch as nine shifted TMA boxes), bias + ReLU epilogues, both branch maps
            kept; `aitb_sk_combine` squares and sums them
  backward  relu(z)**2 is C1 (derivative 2 relu(z)): d_branch = 2 dv r, no mask.  dgrad = the same grouped GEMM with
            the per-group transposed (3x3: tap-flipped) weights, the 1x1 branch accumulated onto the 3x3 one;
            wgrad = ONE `aitb_wgrad_conv` launch per branch (n-tile = group; the 3x3 input read through a 4-D TMA view, one
            shifted box per tap, zero-filled outside the map -- no im2col buffer); bias gradients = column sums

No CPU / eager fallback.
"""
import ctypes as C

import torch

from . import _lib as L
from . import ops
from .packing import round_to_tf32

GROUPS = 8
GC = 128                     # channels per group (1024 / 8)


def _call(fn, *args):
    L.check(fn(*args, L.stream_ptr()))


def _tap_major(w):
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)


def _grouped(x, w, out, G, taps, flags=0, bias=None, round_tf32=False):
    return ops.gemm(x, w, out, M=G * 64, N=1024, K=GC, block_n=128, view="map", map_args=(1024, 8, 8, 1, G), taps=taps,
                    group_c=GC, flags=flags, bias=bias, round_tf32=round_tf32)


def dgrad_weights(m1, m3):
    """Packed forward weights (tap-major, [1024 out, taps*128 in-of-group]) -> the weights of the input-gradient convolution,
    in the same packed form: within each group the in / out roles swap, and the 3x3 taps flip (the adjoint of a 'same'
    correlation is the correlation with the 180-degree rotated kernel).  Pure tensor re-indexing (host logic; CPU-testable)."""
    w3t = m3.view(GROUPS, GC, 9, GC).flip(2).permute(0, 3, 2, 1).contiguous().view(GROUPS * GC, 9 * GC)
    w1t = m1.view(GROUPS, GC, GC).transpose(1, 2).contiguous().view(GROUPS * GC, GC)
    return w1t, w3t


class _SKBlockFn(torch.autograd.Function):
    """forward(x [G,1024,8,8], w1 [1024,128,1,1], b1 [1024], w3 [1024,128,3,3], b3 [1024], cl_out) -> [G,1024,8,8], or --
    cl_out: the consumer is `top_train` -- the channels-last token-major map [G,64,1024] rounded to tf32 (its GEMM operand;
    saves the NCHW round trip in both directions: the incoming gradient then is channels-last too)."""

    @staticmethod
    @L.on_tensor_device
    def forward(ctx, x_nchw, w1, b1, w3, b3, cl_out=False, cl_in=False):
        ops._need_cuda(x_nchw, w1, b1, w3, b3)
        if tuple(x_nchw.shape[1:]) != ((64, 1024) if cl_in else (1024, 8, 8)) or tuple(w1.shape) != (1024, GC, 1, 1) or tuple(w3.shape) != (1024, GC, 3, 3):
            raise RuntimeError("sk_block_train: expected x [G,1024,8,8] and the grouped (groups=8) 1x1 / 3x3 convolution weights")
        lib = L.load()
        G = x_nchw.shape[0]
        dev = x_nchw.device
        M = G * 64
        if cl_in:   # the token-major, tf32-rounded map the AIT training forward left in its saved buffer
            x0 = x_nchw.detach().contiguous().float().view(M, 1024)
        else:
            x0 = ops.transpose_cs(x_nchw.detach().contiguous().float().reshape(G, 1024, 64), True, out_dtype=torch.float32,
                                  round_tf32=True).view(M, 1024)
        m1 = round_to_tf32(_tap_major(w1.detach().float()).contiguous())          # [1024, 128]
        m3 = round_to_tf32(_tap_major(w3.detach().float()).contiguous())          # [1024, 9*128] tap-major
        r1 = torch.empty((M, 1024), dtype=torch.float32, device=dev)
        r3 = torch.empty((M, 1024), dtype=torch.float32, device=dev)
        fl = L.EPI_BIAS | L.EPI_RELU
        _grouped(x0, m1, r1, G, 1, fl, b1.detach().float().contiguous())
        _grouped(x0, m3, r3, G, 9, fl, b3.detach().float().contiguous())
        v = torch.empty((M, 1024), dtype=torch.float32, device=dev)
        _call(lib.aitb_sk_combine, L.ptr(r1), L.ptr(r3), L.ptr(v), C.c_size_t(v.numel()), 1 if cl_out else 0)
        ctx.G, ctx.x0, ctx.m1, ctx.m3, ctx.r1, ctx.r3, ctx.cl_out, ctx.cl_in = G, x0, m1, m3, r1, r3, cl_out, cl_in
        if cl_out:
            return v.view(G, 64, 1024)
        return ops.transpose_cs(v.view(G, 64, 1024), False, out_dtype=torch.float32).view(G, 1024, 8, 8)

    @staticmethod
    @L.on_tensor_device
    def backward(ctx, d_out):
        lib = L.load()
        G, x0, m1, m3, r1, r3 = ctx.G, ctx.x0, ctx.m1, ctx.m3, ctx.r1, ctx.r3
        dev = x0.device
        M = G * 64
        if ctx.cl_out:
            dv = d_out.contiguous().float().view(M, 1024)
        else:
            dv = ops.transpose_cs(d_out.contiguous().float().reshape(G, 1024, 64), True, out_dtype=torch.float32).view(M, 1024)
        d1 = torch.empty_like(dv)
        d3 = torch.empty_like(dv)
        _call(lib.aitb_sk_combine_bwd, L.ptr(dv), L.ptr(r1), L.ptr(r3), L.ptr(d1), L.ptr(d3), C.c_size_t(dv.numel()))
        need_x, need_w1, need_b1, need_w3, need_b3 = ctx.needs_input_grad[:5]
        db1 = ops.colsum(d1) if need_b1 else None
        db3 = ops.colsum(d3) if need_b3 else None
        dw1 = dw3 = None
        # one launch per branch for all eight groups: n-tile = group, X read through the (shifted) TMA map view -- no im2col
        if need_w1:
            dw1 = ops.wgrad_conv(d1, x0, G, 8, 1024, 1024, groups=GROUPS, taps=1)
            dw1 = dw1.view(1024, 1, 1, GC).permute(0, 3, 1, 2).contiguous()
        if need_w3:
            dw3 = ops.wgrad_conv(d3, x0, G, 8, 1024, 1024, groups=GROUPS, taps=9)
            dw3 = dw3.view(1024, 3, 3, GC).permute(0, 3, 1, 2).contiguous()
        dx_nchw = None
        if need_x:
            # per group: dx[:, in] = sum_taps shift(d3)[:, out] W3[out, flipped tap, in] + d1[:, out] W1[out, in]
            w1t, w3t = dgrad_weights(m1, m3)
            dx = torch.empty((M, 1024), dtype=torch.float32, device=dev)
            _grouped(d3, w3t, dx, G, 9)
            _grouped(d1, w1t, dx, G, 1, L.EPI_ACCUM, round_tf32=ctx.cl_in)   # cl_in: the consumer is the AIT backward's GEMMs
            if ctx.cl_in:
                dx_nchw = dx.view(G, 64, 1024)
            else:
                dx_nchw = ops.transpose_cs(dx.view(G, 64, 1024), False, out_dtype=torch.float32).view(G, 1024, 8, 8)
        return dx_nchw, dw1, db1, dw3, db3, None, None


def sk_block_train(blk, x, channels_last_out=False, channels_last_in=False):
    """Differentiable `SKBlock.forward(x)` (blocks_...sk.py:960-984): x [G,1024,8,8] -> [G,1024,8,8]
    (channels_last_out: [G,64,1024] token-major, tf32-rounded -- only for `top_train.head_to_tail_train(..., channels_last=True)`;
    channels_last_in: x is such a map, from `Transformer(..., token_major_out=True)`)."""
    c1, c3 = blk.convs[0][0], blk.convs[1][0]
    return _SKBlockFn.apply(x, c1.weight, c1.bias, c3.weight, c3.bias, bool(channels_last_out), bool(channels_last_in))


def sknet_train(sk, x_props, x_query, channels_last_out=False, channels_last_in=False):
    """Differentiable `SKNet.forward(x_props, x_query)` (blocks_...sk.py:993-998) -> (f_props, f_query).
    channels_last_in applies to x_props only (x_query is the detector's NCHW query feature)."""
    return (sk_block_train(sk.sk_props, x_props, channels_last_out, channels_last_in),
            sk_block_train(sk.sk_query, x_query, channels_last_out))
