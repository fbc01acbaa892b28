"""GPU parity of the training path (BASELINE config 4): backward building blocks and the AIT
forward + backward against the CPU oracle's autograd (fp64) and the reference's own gradients
(tests/golden/ait_grad.pt, generated from the unmodified reference modules)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _tf32(x):
    from ait_b200.packing import round_to_tf32
    return round_to_tf32(x.float())


@pytest.mark.parametrize("M,N,K", [(64, 128, 64), (1000, 512, 64), (4096 + 37, 1536, 512), (777, 512, 1024),
                                   (20000, 2048, 512), (333, 512, 2048), (31, 128, 256)])
def test_wgrad_mn_major_tcgen05(M, N, K):
    """dW = dY^T X with both operands read MN-major straight from the row-major activations."""
    from ait_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    dy = _tf32(torch.randn(M, N, generator=g))
    x = _tf32(torch.randn(M, K, generator=g))
    ref = dy.double().t() @ x.double()
    out = ops.wgrad(dy.to(DEV), x.to(DEV))
    err = float((out.cpu().double() - ref).abs().max() / ref.abs().max())
    assert err < 2e-5, err
    # accumulation into an existing gradient + strided operands (a column block of a wider buffer)
    wide = _tf32(torch.randn(M, N + 256, generator=g)).to(DEV)
    base = torch.ones(N, K, device=DEV)
    out2 = ops.wgrad(wide[:, 128:], x.to(DEV), dw=base.clone(), N=N)
    ref2 = 1.0 + wide[:, 128:128 + N].cpu().double().t() @ x.double()
    assert float((out2.cpu().double() - ref2).abs().max() / ref2.abs().max()) < 2e-5
