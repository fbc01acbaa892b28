import os, sys, json
sys.path.insert(0, "/root/repo")
import torch
from ait_b200 import _lib as L, ops
dev = "cuda:0"
M, N, K = 153600, 1024, 1280
for mode in ("fp32", "bf16"):
    split = mode == "fp32"
    conv = (lambda x: ops.split_planes(x)) if split else (lambda x: x.to(torch.bfloat16))
    a, w = conv(torch.randn(M, K, device=dev)), conv(torch.randn(N, K, device=dev) / 30)
    out = torch.empty(M, N * (2 if split else 1), device=dev, dtype=torch.bfloat16)
    bias = torch.zeros(N, device=dev)
    def t(fn, n=10):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n): fn()
        e.record(); torch.cuda.synchronize()
        return s.elapsed_time(e) / n
    for bn in (128, 256):
        ms = t(lambda: ops.gemm(a, w, out, M=M, N=N, K=K, block_n=bn, flags=L.EPI_BIAS | L.EPI_RELU, bias=bias, split=split))
        print(mode, "plain K=1280 block_n", bn, "2cta" if (bn == 256 and not os.environ.get("AITB_NO_2CTA")) else "1cta", round(ms, 3), "ms", round(2 * M * N * K / ms / 1e9), "TF/s")
