"""A/B of the plain GEMM epilogue's bias prefetch (AITB_NO_BIAS_AHEAD=1 switches it off per call) on the FFN w_1 shape
(M = 153600, N = 2048, K = 512, bias + ReLU), same process, alternating, CUDA events.   python tools/epi_ab.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ait_b200 import _lib as L, ops  # noqa: E402

dev = "cuda:0"
M, N, K = 153600, 2048, 512
for mode in ("bf16", "tf32", "fp32"):
    split = mode == "fp32"
    a = torch.randn(M, K, device=dev)
    w = torch.randn(N, K, device=dev) / K ** 0.5
    if split:
        a, w = ops.split_planes(a), ops.split_planes(w)
        o = torch.empty(M, 2 * N, device=dev, dtype=torch.bfloat16)
    elif mode == "bf16":
        a, w = a.bfloat16(), w.bfloat16()
        o = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    else:
        o = torch.empty(M, N, device=dev)
    bias = torch.zeros(N, device=dev)

    def run():
        ops.gemm(a, w, o, M=M, N=N, K=K, block_n=256, flags=L.EPI_BIAS | L.EPI_RELU, bias=bias, split=split)

    def timeit(reps=20):
        run()
        torch.cuda.synchronize()
        st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st.record()
        for _ in range(reps):
            run()
        en.record()
        torch.cuda.synchronize()
        return st.elapsed_time(en) / reps * 1e3

    res = {"ahead": [], "at_use": []}
    for _ in range(3):
        os.environ.pop("AITB_NO_BIAS_AHEAD", None)
        res["ahead"].append(timeit())
        os.environ["AITB_NO_BIAS_AHEAD"] = "1"
        res["at_use"].append(timeit())
    os.environ.pop("AITB_NO_BIAS_AHEAD", None)
    print(mode, {k: ["%.1f" % x for x in v] for k, v in res.items()}, "us")
    del a, w, o
