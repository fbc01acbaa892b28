"""Signed relative error of the tensor-core accumulation (run on the GPU box): is it biased (RZ) or centred (RN)?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ait_b200 import ops

dev = "cuda:0"
g = torch.Generator().manual_seed(0)
for K in (64, 512, 2048, 4608 if False else 4096):
    M, N = 256, 256
    for kind in ("positive", "signed"):
        a = torch.rand(M, K, generator=g) + 0.5 if kind == "positive" else torch.randn(M, K, generator=g)
        w = (torch.rand(N, K, generator=g) + 0.5 if kind == "positive" else torch.randn(N, K, generator=g)) / K ** 0.5
        # bf16-exact operands -> single-pass bf16 GEMM has NO operand error, only accumulation error
        ab, wb = a.bfloat16(), w.bfloat16()
        ref = ab.double() @ wb.double().t()
        out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        # fp32 output needed: use the split kernel with lo planes = 0 (3 passes, two of them adding zeros)
        a2 = torch.cat([ab, torch.zeros_like(ab)], 1).to(dev)
        w2 = torch.cat([wb, torch.zeros_like(wb)], 1).to(dev)
        o2 = torch.empty(M, 2 * N, dtype=torch.bfloat16, device=dev)
        ops.gemm(a2, w2, o2, M=M, N=N, K=K, block_n=256, split=True)
        res = ops.join_planes(o2).cpu().double()
        rel = (res - ref) / ref.abs().clamp_min(1e-30)
        scale = ref.abs().mean()
        err = (res - ref) * torch.sign(ref) / scale
        print("K=%5d %-8s zero-lo : mean signed err/scale %+.3e  rms %.3e   (n_steps=%d)" % (K, kind, float(err.mean()), float(err.pow(2).mean().sqrt()), 3 * K // 16))
        # full split of fp32 data
        o3 = torch.empty(M, 2 * N, dtype=torch.bfloat16, device=dev)
        ops.gemm(ops.split_planes(a).to(dev), ops.split_planes(w).to(dev), o3, M=M, N=N, K=K, block_n=256, split=True)
        ref3 = a.double() @ w.double().t()
        res3 = ops.join_planes(o3).cpu().double()
        err3 = (res3 - ref3) * torch.sign(ref3) / ref3.abs().mean()
        print("K=%5d %-8s split    : mean signed err/scale %+.3e  rms %.3e" % (K, kind, float(err3.mean()), float(err3.pow(2).mean().sqrt())))
