"""Drop-in for the reference's co-attention block (`CoAttention`,
lib/model/modules/blocks_coatt_transformer_sk.py:17-122) and its wrapper `CoAttentionModule`
(lib/model/faster_rcnn/faster_rcnn_coatt_transformer_sk.py:104-160) in the configuration the detector builds:
in_ch 1024, c_hidden 512, with_residual, 'division' normalisation.  Same parameter names (emb, rho, phi,
omega.{0,1}, theta.{0,1}); executed by `aitb_coattention_forward` -- tcgen05 GEMMs for the five 1x1 convolutions
and the three attention products (the contraction over image positions runs on the MN-major kernel, no
transposed copy of the map), GroupNorm + residual kernels -- in fp32 storage with tf32 tensor-core math.
`.train()` with gradients enabled runs the differentiable step of ait_b200/coatt_train.py (own backward).
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from .packing import fingerprint, round_to_tf32


class CoattWeights(C.Structure):
    _fields_ = [("dtype", C.c_int), ("round_tf32", C.c_int), ("w_emb_phi", C.c_void_p), ("b_emb_phi", C.c_void_p),
                ("emb", L.Linear), ("rho", L.Linear), ("theta", L.Linear), ("omega", L.Linear),
                ("theta_gn", L.LNorm), ("omega_gn", L.LNorm)]


class CoAttention(nn.Module):
    def __init__(self, **kwargs):
        super().__init__()
        self.in_ch = kwargs.get("in_ch", 1024)
        self.c_hidden = kwargs.get("c_hidden", 512)
        self.with_residual = kwargs.get("with_residual", True)
        self.normlization = kwargs.get("normlization", "division")
        if (self.in_ch, self.c_hidden, self.with_residual, self.normlization) != (1024, 512, True, "division"):
            raise NotImplementedError("ait_b200.CoAttention supports the detector's configuration only: in_ch=1024, "
                                      "c_hidden=512, with_residual=True, normlization='division'")
        self.emb = nn.Conv2d(self.in_ch, self.c_hidden, kernel_size=1)
        self.rho = nn.Conv2d(self.in_ch, self.c_hidden, kernel_size=1)
        self.phi = nn.Conv2d(self.in_ch, self.c_hidden, kernel_size=1)
        self.omega = nn.Sequential(nn.Conv2d(self.c_hidden, self.in_ch, kernel_size=1), nn.GroupNorm(32, self.in_ch))
        self.theta = nn.Sequential(nn.Conv2d(self.c_hidden, self.in_ch, kernel_size=1), nn.GroupNorm(32, self.in_ch))
        for m in self.modules():                       # reset_params (:50-58): the block starts as the identity
            if isinstance(m, nn.GroupNorm):
                nn.init.constant_(m.weight, 0)
                nn.init.constant_(m.bias, 0)
        self._packed = None

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def _load_from_state_dict(self, *a, **k):
        super()._load_from_state_dict(*a, **k)
        self._packed = None

    def invalidate(self):
        self._packed = None

    def _pack(self):
        def mat(t):
            return round_to_tf32(t.detach().float().flatten(1).contiguous())

        def vec(t):
            return t.detach().float().contiguous()

        with torch.no_grad():
            keep = dict(
                w_emb_phi=mat(torch.cat([self.emb.weight, self.phi.weight], 0)),
                b_emb_phi=vec(torch.cat([self.emb.bias, self.phi.bias], 0)),
                emb_w=mat(self.emb.weight), emb_b=vec(self.emb.bias), rho_w=mat(self.rho.weight), rho_b=vec(self.rho.bias),
                theta_w=mat(self.theta[0].weight), theta_b=vec(self.theta[0].bias),
                omega_w=mat(self.omega[0].weight), omega_b=vec(self.omega[0].bias),
                theta_g=vec(self.theta[1].weight), theta_be=vec(self.theta[1].bias),
                omega_g=vec(self.omega[1].weight), omega_be=vec(self.omega[1].bias))
        w = CoattWeights()
        w.dtype, w.round_tf32 = L.AITB_F32, 1
        w.w_emb_phi, w.b_emb_phi = keep["w_emb_phi"].data_ptr(), keep["b_emb_phi"].data_ptr()
        w.emb.w, w.emb.bias = keep["emb_w"].data_ptr(), keep["emb_b"].data_ptr()
        w.rho.w, w.rho.bias = keep["rho_w"].data_ptr(), keep["rho_b"].data_ptr()
        w.theta.w, w.theta.bias = keep["theta_w"].data_ptr(), keep["theta_b"].data_ptr()
        w.omega.w, w.omega.bias = keep["omega_w"].data_ptr(), keep["omega_b"].data_ptr()
        w.theta_gn.gamma, w.theta_gn.beta = keep["theta_g"].data_ptr(), keep["theta_be"].data_ptr()
        w.omega_gn.gamma, w.omega_gn.beta = keep["omega_g"].data_ptr(), keep["omega_be"].data_ptr()
        self._packed = (w, keep)
        return w

    def forward(self, x_img, x_qry):
        """x_img [B,1024,H,W], x_qry [B,1024,8,8] -> (non_img [B,1024,H,W], non_qry [B,1024,8,8])."""
        if self.training and torch.is_grad_enabled() and (
                x_img.requires_grad or x_qry.requires_grad or any(p.requires_grad for p in self.parameters())):
            from . import coatt_train          # differentiable training step (forward keeping activations + own backward)
            return coatt_train.coattention_train(self, x_img, x_qry)
        lib = L.load()
        ops._need_cuda(x_img, x_qry)
        if x_img.dtype != torch.float32 or x_qry.dtype != torch.float32:
            raise RuntimeError("ait_b200.CoAttention: float32 feature maps required")
        B, c, H, W = x_img.shape
        if c != 1024 or tuple(x_qry.shape) != (B, 1024, 8, 8):
            raise RuntimeError("ait_b200.CoAttention: expected x_img [B,1024,H,W] and x_qry [B,1024,8,8]")
        fp = fingerprint(self)                  # in-place parameter updates since the last packing?
        if self._packed is None or self._packed_fp != fp:
            self._packed_fp = fp
            self._pack()
        w = self._packed[0]
        x_img, x_qry = x_img.contiguous(), x_qry.contiguous()
        non_img, non_qry = torch.empty_like(x_img), torch.empty_like(x_qry)
        nbytes = lib.aitb_coattention_workspace_bytes(B, H, W)
        ws = ops._workspace(nbytes, x_img.device, "coatt")
        with torch.cuda.device(x_img.device):
            L.check(lib.aitb_coattention_forward(C.byref(w), L.ptr(x_img), L.ptr(x_qry), B, H, W, L.ptr(non_img),
                                                 L.ptr(non_qry), L.ptr(ws), nbytes, L.stream_ptr()))
        return non_img, non_qry


class CoAttentionModule(nn.Module):
    """faster_rcnn_coatt_transformer_sk.py:104-160: `self.coattention = B.CoAttention(in_ch, in_ch // 2, residual, division)`."""

    def __init__(self, inplanes):
        super().__init__()
        self.in_ch = inplanes
        self.c_hidden = max(inplanes // 2, 1)
        self.coattention = CoAttention(in_ch=self.in_ch, c_hidden=self.c_hidden, with_residual=True,
                                       normlization="division")

    def forward(self, x_img, x_qry):
        return self.coattention(x_img, x_qry)
