"""Launch the three hot kernels in isolation on benchmark-shaped data (for `ncu --set full`):
FFN w_1 GEMM, encoder self-attention core, ROIAlign (token-major)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ait_b200 import _lib as L, ops, synth  # noqa: E402
from ait_b200.proposal import propose_rois  # noqa: E402

dev = "cuda:0"
B, P = 8, 300
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
# ROIAlign inputs
maps = torch.stack([synth.c4_map(u) for u in range(B)]).to(dev)
rpn = [synth.rpn_outputs(u) for u in range(B)]
rois, _ = propose_rois(torch.stack([r[0] for r in rpn]).to(dev), torch.stack([r[1] for r in rpn]).to(dev))
nhwc = ops.transpose_cs(maps.reshape(B, 1024, -1), True).view(B, 38, 63, 1024)
# GEMM inputs (FFN w_1)
M, N, K = B * P * 64, 2048, 512
a = torch.randn(M, K, device=dev)
w = torch.randn(N, K, device=dev) / K ** 0.5
o = torch.empty(M, N, device=dev)
bias = torch.zeros(N, device=dev)
# attention inputs
G = B * P
qkv = torch.randn(G * 64, 1536, device=dev)
w_sk = torch.randn(512, 64, device=dev) * 0.1
b_sk = torch.zeros(512, device=dev)
ao = torch.empty(G, 64, 64, device=dev)
torch.cuda.synchronize()
for _ in range(reps):
    ops.gemm(a, w, o, M=M, N=N, K=K, block_n=256, flags=L.EPI_BIAS | L.EPI_RELU, bias=bias)
    ops.attn_core(qkv, 1536, 1, qkv.view(-1)[512:], qkv.view(-1)[1024:], 1536, w_sk, b_sk, G, 0, 49, ao)
    ops.roi_align_forward(nhwc, rois.view(-1, 5), 1 / 16.0, 7, 7, 0, token_major=True)
torch.cuda.synchronize()
print("done")
