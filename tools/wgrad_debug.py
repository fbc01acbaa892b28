"""Probe the MN-major operand layout of the wgrad kernel with structured inputs."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ait_b200 import ops  # noqa: E402

dev = "cuda:0"
M, N, K = 32, 128, 64
torch.set_printoptions(linewidth=200)
# 1. all-ones dy: out[n, k] = sum_m x[m, k]
dy = torch.ones(M, N)
x = torch.arange(K).float()[None, :].repeat(M, 1)          # x[m, k] = k  -> expect out[n, k] = 32 k
out = ops.wgrad(dy.to(dev), x.to(dev)).cpu()
print("probe1 expect 32*k per column; row0:", out[0, :16].tolist())
print("        row 77:", out[77, :16].tolist(), " distinct rows:", len({tuple(r.tolist()) for r in out}))
# 2. x all ones, dy[m, n] = n -> out[n, k] = 32 n
dy = torch.arange(N).float()[None, :].repeat(M, 1)
x = torch.ones(M, K)
out = ops.wgrad(dy.to(dev), x.to(dev)).cpu()
print("probe2 expect 32*n per row; col0:", out[:16, 0].tolist(), "...", out[120:, 0].tolist())
# 3. row dependence: dy[m, n] = (m == m0), x = 1 -> out = 1 everywhere for every m0
for m0 in (0, 1, 7, 8, 31):
    dy = torch.zeros(M, N)
    dy[m0] = 1.0
    out = ops.wgrad(dy.to(dev), torch.ones(M, K).to(dev)).cpu()
    print("probe3 m0=%d: min %.1f max %.1f sum %.1f (expect all 1, sum %d)" % (m0, out.min(), out.max(), out.sum(), N * K))
# 4. pairing of rows between A and B: dy[m, :] = (m == m0), x[m, :] = (m == m1) -> out = (m0 == m1)
for m0, m1 in ((3, 3), (3, 4), (12, 12), (12, 20)):
    dy = torch.zeros(M, N)
    dy[m0] = 1.0
    x = torch.zeros(M, K)
    x[m1] = 1.0
    out = ops.wgrad(dy.to(dev), x.to(dev)).cpu()
    print("probe4 m0=%d m1=%d: sum %.1f (expect %d)" % (m0, m1, out.sum(), N * K if m0 == m1 else 0))
