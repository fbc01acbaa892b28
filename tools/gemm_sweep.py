"""Time the plain bn256 GEMM on benchmark-sized shapes (run twice: default and AITB_NO_2CTA=1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ait_b200 import _lib as L, ops

dev = "cuda:0"
M = 153600
for dtype in (torch.float32, torch.bfloat16):
    for (N, K, flags) in [(2048, 512, L.EPI_BIAS | L.EPI_RELU), (2048, 512, 0), (512, 2048, 0), (1536, 512, 0), (1024, 512, L.EPI_BIAS),
                          (512, 4096, 0), (256, 8192, 0)]:
        a = torch.randn(M, K, device=dev).to(dtype)
        w = (torch.randn(N, K, device=dev) / K ** 0.5).to(dtype)
        o = torch.empty(M, N, device=dev, dtype=dtype)
        bias = torch.zeros(N, device=dev)
        f = lambda: ops.gemm(a, w, o, M=M, N=N, K=K, block_n=256, flags=flags, bias=bias if flags else None)
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st.record()
        for _ in range(10):
            f()
        en.record()
        torch.cuda.synchronize()
        ms = st.elapsed_time(en) / 10
        print("%s N=%5d K=%5d flags=%3d  %.3f ms  %.0f TFLOP/s" % (str(dtype)[6:], N, K, flags, ms, 2.0 * M * N * K / ms / 1e9))
        del a, w, o
