"""The co-attention block (`CoAttention.forward`, lib/model/modules/blocks_coatt_transformer_sk.py:60-122; 'division'
normalisation, with_residual) as a differentiable function on the device -- row f3's training path.  The reference obtains
the backward from torch autograd; here forward and backward are composed from the library's building blocks in fp32
storage with tf32 tensor-core math, like the rest of the training step:

  forward   the five 1x1 convolutions as tcgen05 GEMMs on token-major maps ([emb | phi] of the image in one launch), the
            three attention products per image (co^T = phi_img rho_qry^T, non_img = co^T emb_qry / N_q, non_qry^T =
            emb_img^T co^T on the MN-major row-contraction kernel), GroupNorm(32) + identity (`aitb_group_norm_forward`,
            which keeps the per-(image, group) sums); every activation the backward needs is kept
  backward  GroupNorm backward (`aitb_group_norm_backward`), dgrad = the same GEMM kernel with transposed weights, wgrad =
            `aitb_wgrad` (dW += dY^T X on MN-major operands), bias gradients = `aitb_colsum`; the attention products are
            differentiated product by product (the 1 / N_q and 1 / N_i of the 'division' normalisation ride in the GEMMs'
            out_scale)

Gradients are returned for both inputs and all 14 parameters (emb, rho, phi, omega.{0,1}, theta.{0,1}: weight and bias).
No CPU / eager fallback.
"""
import torch

from . import _lib as L
from . import ops
from .packing import round_to_tf32

PARAM_NAMES = ["emb.weight", "emb.bias", "rho.weight", "rho.bias", "phi.weight", "phi.bias",
               "omega.0.weight", "omega.0.bias", "omega.1.weight", "omega.1.bias",
               "theta.0.weight", "theta.0.bias", "theta.1.weight", "theta.1.bias"]


def _mat(w):
    return round_to_tf32(w.detach().float().flatten(1).contiguous())


def _vec(b):
    return b.detach().float().contiguous()


def _empty(rows, cols, dev):
    return torch.empty((rows, cols), dtype=torch.float32, device=dev)


def _gn_forward(x, identity, gamma, beta, B, N):
    lib = L.load()
    sums = torch.empty((B, 32, 2), dtype=torch.float64, device=x.device)
    out = torch.empty_like(x)
    L.check(lib.aitb_group_norm_forward(L.ptr(x), L.ptr(identity), L.ptr(gamma), L.ptr(beta), B, N, 32, 1e-5, L.ptr(sums),
                                        L.ptr(out), L.stream_ptr()))
    return out, sums


def _gn_backward(dy, x, sums, gamma, B, N):
    """-> (dx rounded to tf32, dgamma, dbeta)"""
    lib = L.load()
    dev = x.device
    bsums = torch.empty((B, 32, 2), dtype=torch.float64, device=dev)
    dx = torch.empty_like(x)
    dgamma = torch.zeros(1024, dtype=torch.float32, device=dev)
    dbeta = torch.zeros(1024, dtype=torch.float32, device=dev)
    L.check(lib.aitb_group_norm_backward(L.ptr(dy), L.ptr(x), L.ptr(sums), L.ptr(gamma), B, N, 32, 1e-5, 1, L.ptr(bsums),
                                         L.ptr(dx), L.ptr(dgamma), L.ptr(dbeta), L.stream_ptr()))
    return dx, dgamma, dbeta


class _CoAttentionFn(torch.autograd.Function):
    """forward(x_img [B,1024,H,W], x_qry [B,1024,8,8], *params in PARAM_NAMES order) -> (non_img, non_qry)."""

    @staticmethod
    @L.on_tensor_device
    def forward(ctx, x_img, x_qry, *params):
        ops._need_cuda(x_img, x_qry, *params)
        B, c, H, W = x_img.shape
        if c != 1024 or tuple(x_qry.shape) != (B, 1024, 8, 8) or len(params) != 14:
            raise RuntimeError("coattention_train: expected x_img [B,1024,H,W], x_qry [B,1024,8,8] and the 14 parameters")
        p = dict(zip(PARAM_NAMES, params))
        dev = x_img.device
        Ni, rows, rq = H * W, B * H * W, B * 64
        w_ep = _mat(torch.cat([p["emb.weight"], p["phi.weight"]], 0))            # [1024, 1024]: rows 0..511 emb, 512.. phi
        b_ep = _vec(torch.cat([p["emb.bias"], p["phi.bias"]], 0))
        w_emb, w_rho = w_ep[:512].contiguous(), _mat(p["rho.weight"])
        w_th, w_om = _mat(p["theta.0.weight"]), _mat(p["omega.0.weight"])        # [1024, 512]
        xi3 = x_img.detach().float().contiguous().view(B, 1024, Ni)
        xq3 = x_qry.detach().float().contiguous().view(B, 1024, 64)
        Xi = ops.transpose_cs(xi3, True).view(rows, 1024)                        # exact tokens (residual)
        Xir = ops.transpose_cs(xi3, True, round_tf32=True).view(rows, 1024)      # MMA operand
        Xq = ops.transpose_cs(xq3, True).view(rq, 1024)
        Xqr = ops.transpose_cs(xq3, True, round_tf32=True).view(rq, 1024)
        EP = ops.gemm(Xir, w_ep, _empty(rows, 1024, dev), M=rows, N=1024, K=1024, block_n=256, flags=L.EPI_BIAS, bias=b_ep,
                      round_tf32=True)                                           # [emb | phi] of the image (:70-79)
        Eq = ops.gemm(Xqr, w_emb, _empty(rq, 512, dev), M=rq, N=512, K=1024, block_n=256, flags=L.EPI_BIAS,
                      bias=_vec(p["emb.bias"]), round_tf32=True)
        Rq = ops.gemm(Xqr, w_rho, _empty(rq, 512, dev), M=rq, N=512, K=1024, block_n=256, flags=L.EPI_BIAS,
                      bias=_vec(p["rho.bias"]), round_tf32=True)
        EqT = ops.transpose_cs(Eq.view(B, 64, 512), True)                        # [B, 512, 64]
        coT = _empty(rows, 64, dev)
        NI = _empty(rows, 512, dev)
        NQT = torch.zeros((B, 512, 64), dtype=torch.float32, device=dev)
        for i in range(B):
            EPb, coTb = EP[i * Ni:(i + 1) * Ni], coT[i * Ni:(i + 1) * Ni]
            # co^T = phi_img rho_qry^T  [Ni, 64]  (:81)
            ops.gemm(EPb[:, 512:], Rq[i * 64:(i + 1) * 64], coTb, M=Ni, N=64, K=512, block_n=64, lda=1024, round_tf32=True)
            # non_img = (co^T / N_q) emb_qry  [Ni, 512]  (:92-93, 99)
            ops.gemm(coTb, EqT[i], NI[i * Ni:(i + 1) * Ni], M=Ni, N=512, K=64, block_n=256, out_scale=1.0 / 64.0,
                     round_tf32=True)
            # (non_qry * N_i)^T = emb_img^T co^T  [512, 64]: contraction over the image positions (:104)
            ops.wgrad(EPb, coTb, dw=NQT[i], N=512, K=64)
        TH = ops.gemm(NI, w_th, _empty(rows, 1024, dev), M=rows, N=1024, K=512, block_n=256, flags=L.EPI_BIAS,
                      bias=_vec(p["theta.0.bias"]))
        g_th, g_om = _vec(p["theta.1.weight"]), _vec(p["omega.1.weight"])
        OUTi, sums_th = _gn_forward(TH, Xi, g_th, _vec(p["theta.1.bias"]), B, Ni)
        non_img = ops.transpose_cs(OUTi.view(B, Ni, 1024), False).view(B, 1024, H, W)
        NQ = ops.transpose_cs(NQT, True, round_tf32=True).view(rq, 512)          # [B, 64, 512], still times N_i
        OM = ops.gemm(NQ, w_om, _empty(rq, 1024, dev), M=rq, N=1024, K=512, block_n=256, flags=L.EPI_BIAS,
                      bias=_vec(p["omega.0.bias"]), out_scale=1.0 / Ni)
        OUTq, sums_om = _gn_forward(OM, Xq, g_om, _vec(p["omega.1.bias"]), B, 64)
        non_qry = ops.transpose_cs(OUTq.view(B, 64, 1024), False).view(B, 1024, 8, 8)
        ctx.dims = (B, H, W)
        ctx.keep = dict(Xir=Xir, Xqr=Xqr, EP=EP, Eq=Eq, Rq=Rq, coT=coT, NI=NI, NQ=NQ, TH=TH, OM=OM, sums_th=sums_th,
                        sums_om=sums_om, w_ep=w_ep, w_rho=w_rho, w_th=w_th, w_om=w_om, g_th=g_th, g_om=g_om)
        ctx.wshapes = [tuple(t.shape) for t in params]
        ctx.in_dtypes = (x_img.dtype, x_qry.dtype)
        return non_img.to(x_img.dtype), non_qry.to(x_qry.dtype)

    @staticmethod
    @L.on_tensor_device
    def backward(ctx, g_img, g_qry):
        if ctx.keep is None:
            raise RuntimeError("ait_b200.CoAttention: backward a second time: the saved activations were freed after the "
                               "first backward (retain_graph=True is not supported; run the forward again)")
        k = ctx.keep
        B, H, W = ctx.dims
        Ni, rows, rq = H * W, B * H * W, B * 64
        dev = k["EP"].device
        T = lambda w: w.t().contiguous()                                         # noqa: E731  [N, K] -> [K, N]
        gOi = ops.transpose_cs(g_img.contiguous().float().view(B, 1024, Ni), True).view(rows, 1024)
        gOq = ops.transpose_cs(g_qry.contiguous().float().view(B, 1024, 64), True).view(rq, 1024)
        # ---- theta: OUTi = GN(TH) + Xi, TH = NI Wth^T + b
        gTH, d_gth, d_bth = _gn_backward(gOi, k["TH"], k["sums_th"], k["g_th"], B, Ni)
        dW_th = ops.wgrad(gTH, k["NI"], N=1024, K=512)
        db_th = ops.colsum(gTH)
        # gradient of NI, already divided by N_q (NI = co^T emb_qry / N_q): both products below then need no scale
        gNI = ops.gemm(gTH, T(k["w_th"]), _empty(rows, 512, dev), M=rows, N=512, K=1024, block_n=256, out_scale=1.0 / 64.0,
                       round_tf32=True)
        # ---- omega: OUTq = GN(OM) + Xq, OM = (NQ / N_i) Wom^T + b  (NQ is kept times N_i)
        gOM, d_gom, d_bom = _gn_backward(gOq, k["OM"], k["sums_om"], k["g_om"], B, 64)
        dW_om = ops.wgrad(gOM, k["NQ"], N=1024, K=512).mul_(1.0 / Ni)
        db_om = ops.colsum(gOM)
        gU = ops.gemm(gOM, T(k["w_om"]), _empty(rq, 512, dev), M=rq, N=512, K=1024, block_n=256, out_scale=1.0 / Ni,
                      round_tf32=True)                                           # gradient of U = co emb_img = NQ * N_i
        gUT = ops.transpose_cs(gU.view(B, 64, 512), True)                        # [B, 512, 64]
        RqT = ops.transpose_cs(k["Rq"].view(B, 64, 512), True)
        gcoT = _empty(rows, 64, dev)
        gEP = _empty(rows, 1024, dev)                                            # [d emb_img | d phi_img]
        gEqT = torch.zeros((B, 512, 64), dtype=torch.float32, device=dev)
        gRqT = torch.zeros((B, 512, 64), dtype=torch.float32, device=dev)
        for i in range(B):
            sl, sq = slice(i * Ni, (i + 1) * Ni), slice(i * 64, (i + 1) * 64)
            EPb, coTb, gNIb, gcoTb, gEPb = k["EP"][sl], k["coT"][sl], gNI[sl], gcoT[sl], gEP[sl]
            # d co^T = gNI emb_qry^T + emb_img gU^T   [Ni, 64]
            ops.gemm(gNIb, k["Eq"][sq], gcoTb, M=Ni, N=64, K=512, block_n=64)
            ops.gemm(EPb, gU[sq], gcoTb, M=Ni, N=64, K=512, block_n=64, lda=1024, flags=L.EPI_ACCUM, round_tf32=True)
            # d emb_qry^T += gNI^T co^T   [512, 64]
            ops.wgrad(gNIb, coTb, dw=gEqT[i], N=512, K=64)
            # d emb_img = co^T gU   [Ni, 512] -> columns 0..511 of gEP
            ops.gemm(coTb, gUT[i], gEPb, M=Ni, N=512, K=64, block_n=256, ldo=1024, round_tf32=True)
            # d phi_img = d co^T rho_qry   [Ni, 512] -> columns 512.. of gEP;  d rho_qry^T = phi_img^T d co^T   [512, 64]
            ops.gemm(gcoTb, RqT[i], gEPb[:, 512:], M=Ni, N=512, K=64, block_n=256, ldo=1024, round_tf32=True)
            ops.wgrad(EPb[:, 512:], gcoTb, dw=gRqT[i], N=512, K=64)
        # ---- the image-side projections [emb | phi]
        dW_ep = ops.wgrad(gEP, k["Xir"], N=1024, K=1024)
        db_ep = ops.colsum(gEP)
        gXi = ops.gemm(gEP, T(k["w_ep"]), _empty(rows, 1024, dev), M=rows, N=1024, K=1024, block_n=256, flags=L.EPI_RES,
                       res=gOi, ldr=1024)                                        # + the identity path
        # ---- the query-side projections emb, rho
        gEq = ops.transpose_cs(gEqT, True, round_tf32=True).view(rq, 512)        # [B, 64, 512]
        gRq = ops.transpose_cs(gRqT, True, round_tf32=True).view(rq, 512)
        dW_emb = dW_ep[:512] + ops.wgrad(gEq, k["Xqr"], N=512, K=1024)
        db_emb = db_ep[:512] + ops.colsum(gEq)
        dW_rho = ops.wgrad(gRq, k["Xqr"], N=512, K=1024)
        db_rho = ops.colsum(gRq)
        gXq = ops.gemm(gEq, T(k["w_ep"][:512]), _empty(rq, 1024, dev), M=rq, N=1024, K=512, block_n=256, flags=L.EPI_RES,
                       res=gOq, ldr=1024)
        ops.gemm(gRq, T(k["w_rho"]), gXq, M=rq, N=1024, K=512, block_n=256, flags=L.EPI_ACCUM)
        d_img = ops.transpose_cs(gXi.view(B, Ni, 1024), False).view(B, 1024, H, W)
        d_qry = ops.transpose_cs(gXq.view(B, 64, 1024), False).view(B, 1024, 8, 8)
        ctx.keep = None
        grads = {"emb.weight": dW_emb, "emb.bias": db_emb, "rho.weight": dW_rho, "rho.bias": db_rho,
                 "phi.weight": dW_ep[512:], "phi.bias": db_ep[512:], "omega.0.weight": dW_om, "omega.0.bias": db_om,
                 "omega.1.weight": d_gom, "omega.1.bias": d_bom, "theta.0.weight": dW_th, "theta.0.bias": db_th,
                 "theta.1.weight": d_gth, "theta.1.bias": d_bth}
        outs = [grads[n].reshape(shp) for n, shp in zip(PARAM_NAMES, ctx.wshapes)]
        return (d_img.to(ctx.in_dtypes[0]), d_qry.to(ctx.in_dtypes[1])) + tuple(outs)


def coattention_train(module, x_img, x_qry):
    """Differentiable `CoAttention.forward(x_img, x_qry)` over the module's own Parameters (the reference's module or
    ait_b200.coattention.CoAttention: same parameter names) -> (non_img, non_qry)."""
    sd = dict(module.named_parameters())
    return _CoAttentionFn.apply(x_img, x_qry, *[sd[n] for n in PARAM_NAMES])
