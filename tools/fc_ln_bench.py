"""fc (64 -> 512) + residual + LayerNorm at the benchmark shape: the streaming kernel (fc_ln.cu) against the tcgen05
GEMM with the LayerNorm epilogue, hot, CUDA events.   python tools/fc_ln_bench.py [fp32|bf16]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ait_b200 import _lib as L, ops  # noqa: E402

dev = "cuda:0"
for mode in (sys.argv[1:] or ["fp32", "bf16"]):
    split = mode == "fp32"
    M = 8 * 300 * 64
    conv = (lambda x: ops.split_planes(x)) if split else (lambda x: x.to(torch.bfloat16))
    a, w = conv(torch.randn(M, 64, device=dev)), conv(torch.randn(512, 64, device=dev) / 8)
    res = conv(torch.randn(M, 512, device=dev))
    gamma, beta = torch.ones(512, device=dev), torch.zeros(512, device=dev)
    out = torch.empty_like(res)
    out2 = torch.empty_like(res)

    def t(fn, n=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / n * 1e3

    t_new = t(lambda: ops.fc_ln(a, w, res, gamma, beta, out, M=M, split=split, res_div=64, res_rep=1))
    t_old = t(lambda: ops.gemm(a, w, out2, M=M, N=512, K=64, block_n=512, flags=L.EPI_RES | L.EPI_LN, res=res, ldr=512,
                               res_div=64, res_rep=1, gamma=gamma, beta=beta, split=split))
    byt = M * 512 * (4 if split else 2) * 2 + M * 64 * (4 if split else 2)
    print(json.dumps({"mode": mode, "fc_ln_us": round(t_new, 1), "gemm_ln_us": round(t_old, 1),
                      "fc_ln_GBs": round(byt / t_new / 1e3, 1), "max_diff": float((out.float() - out2.float()).abs().max())}))
