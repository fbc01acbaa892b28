"""GPU parity of the tcgen05 GEMM building block and the attention core, against fp64 references
computed from tf32-rounded operands (so the only difference left is fp32 accumulation order)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _tf32(x):
    from ait_b200.packing import round_to_tf32
    return round_to_tf32(x.float())


def _prep(x, dtype):
    return _tf32(x) if dtype == torch.float32 else x.to(torch.bfloat16).float()


def _tol(dtype):
    return dict(rtol=2e-5, atol=2e-4) if dtype == torch.float32 else dict(rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,N,K,bn", [(128, 256, 64, 256), (300, 512, 512, 256), (1000, 1536, 512, 256),
                                      (77, 128, 128, 128), (4096 + 5, 2048, 512, 256), (640, 256, 2048, 256)])
def test_gemm_plain_bias_relu(dtype, M, N, K, bn):
    from ait_b200 import _lib as L, ops
    g = torch.Generator().manual_seed(M + N + K)
    a = _prep(torch.randn(M, K, generator=g), dtype)
    w = _prep(torch.randn(N, K, generator=g) / K ** 0.5, dtype)
    bias = torch.randn(N, generator=g)
    ref = F.relu(a.double() @ w.double().t() + bias.double())
    out = torch.full((M, N), float("nan"), dtype=dtype, device=DEV)
    ops.gemm(a.to(DEV, dtype), w.to(DEV, dtype), out, M=M, N=N, K=K, block_n=bn, flags=L.EPI_BIAS | L.EPI_RELU,
             bias=bias.to(DEV))
    torch.testing.assert_close(out.float().cpu().double(), ref, **_tol(dtype))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gemm_layernorm_residual_pos_rowmap(dtype):
    """enc_emb-style epilogue: 49 -> 64 row remap, + bias + pos table, LayerNorm over N = 512."""
    from ait_b200 import _lib as L, ops
    g = torch.Generator().manual_seed(1)
    pairs, K = 11, 1024
    M = pairs * 49
    a = _prep(torch.randn(M, K, generator=g), dtype)
    w = _prep(torch.randn(512, K, generator=g) / K ** 0.5, dtype)
    bias, pos = torch.randn(512, generator=g), torch.randn(64, 512, generator=g)
    gamma, beta = torch.rand(512, generator=g) + 0.5, torch.randn(512, generator=g)
    y = (a.double() @ w.double().t() + bias.double()).view(pairs, 49, 512) + pos[:49].double()
    ref = F.layer_norm(y, (512,), gamma.double(), beta.double(), eps=1e-6)
    out = torch.zeros((pairs * 64, 512), dtype=dtype, device=DEV)
    ops.gemm(a.to(DEV, dtype), w.to(DEV, dtype), out, M=M, N=512, K=K, block_n=512,
             flags=L.EPI_BIAS | L.EPI_POS | L.EPI_LN, bias=bias.to(DEV), pos=pos.to(DEV), pos_rows=64,
             gamma=gamma.to(DEV), beta=beta.to(DEV), rows_in=49, rows_out=64)
    o = out.float().cpu().view(pairs, 64, 512)
    torch.testing.assert_close(o[:, :49].double(), ref, **_tol(dtype))
    assert torch.all(o[:, 49:] == 0)                      # pad rows are not touched by the GEMM


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gemm_residual_broadcast_layernorm(dtype):
    """cross-attention fc epilogue: residual row = (row / 64 / P) * 64 + row % 64."""
    from ait_b200 import _lib as L, ops
    g = torch.Generator().manual_seed(2)
    B, P = 3, 5
    M = B * P * 64
    a = _prep(torch.randn(M, 64, generator=g), dtype)
    w = _prep(torch.randn(512, 64, generator=g) / 8, dtype)
    res = _prep(torch.randn(B * 64, 512, generator=g), dtype)
    gamma, beta = torch.rand(512, generator=g) + 0.5, torch.randn(512, generator=g)
    y = (a.double() @ w.double().t()).view(B, P, 64, 512) + res.double().view(B, 1, 64, 512)
    ref = F.layer_norm(y, (512,), gamma.double(), beta.double(), eps=1e-6).view(M, 512)
    out = torch.zeros((M, 512), dtype=dtype, device=DEV)
    ops.gemm(a.to(DEV, dtype), w.to(DEV, dtype), out, M=M, N=512, K=64, block_n=512, flags=L.EPI_RES | L.EPI_LN,
             res=res.to(DEV, dtype), ldr=512, res_div=64, res_rep=P, gamma=gamma.to(DEV), beta=beta.to(DEV))
    torch.testing.assert_close(out.float().cpu().double(), ref, **_tol(dtype))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("G", [1, 5, 19])
def test_gemm_conv3x3_on_4x4_map_via_shifted_tma(dtype, G):
    """layer4 conv2: 3x3, pad 1, 512 -> 512 on a 4x4 map == 9 shifted TMA boxes with zero fill."""
    from ait_b200 import _lib as L, ops
    from ait_b200.packing import HeadEngine
    g = torch.Generator().manual_seed(G)
    x = _prep(torch.randn(G, 512, 4, 4, generator=g), dtype)
    w = _prep(torch.randn(512, 512, 3, 3, generator=g) / 68.0, dtype)
    bias = torch.randn(512, generator=g)
    ref = F.relu(F.conv2d(x.double(), w.double(), bias.double(), padding=1)).permute(0, 2, 3, 1).reshape(G * 16, 512)
    xt = x.permute(0, 2, 3, 1).contiguous().to(DEV, dtype)
    out = torch.zeros((G * 16, 512), dtype=dtype, device=DEV)
    ops.gemm(xt, HeadEngine._tap_major(w).contiguous().to(DEV, dtype), out, M=G * 16, N=512, K=512, block_n=256,
             view="map", map_args=(512, 4, 4, 1, G), taps=9, flags=L.EPI_BIAS | L.EPI_RELU, bias=bias.to(DEV))
    torch.testing.assert_close(out.float().cpu().double(), ref, **_tol(dtype))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gemm_strided_1x1_and_grouped_convs(dtype):
    from ait_b200 import _lib as L, ops
    from ait_b200.packing import HeadEngine
    g = torch.Generator().manual_seed(4)
    G = 7
    x = _prep(torch.randn(G, 1024, 8, 8, generator=g), dtype)
    xt = x.permute(0, 2, 3, 1).contiguous().to(DEV, dtype)
    # stride-2 1x1 conv 1024 -> 512 (layer4.0.conv1): a strided TMA view, no gather kernel
    w = _prep(torch.randn(512, 1024, 1, 1, generator=g) / 32.0, dtype)
    ref = F.conv2d(x.double(), w.double(), stride=2).permute(0, 2, 3, 1).reshape(G * 16, 512)
    out = torch.zeros((G * 16, 512), dtype=dtype, device=DEV)
    ops.gemm(xt, w.flatten(1).contiguous().to(DEV, dtype), out, M=G * 16, N=512, K=1024, block_n=256, view="map",
             map_args=(1024, 8, 4, 2, G))
    torch.testing.assert_close(out.float().cpu().double(), ref, **_tol(dtype))
    # SKBlock: relu(conv1x1_g8)^2 + relu(conv3x3_g8)^2
    w1 = _prep(torch.randn(1024, 128, 1, 1, generator=g) / 11.0, dtype)
    w3 = _prep(torch.randn(1024, 128, 3, 3, generator=g) / 34.0, dtype)
    b1, b3 = torch.randn(1024, generator=g) * 0.1, torch.randn(1024, generator=g) * 0.1
    f1 = F.relu(F.conv2d(x.double(), w1.double(), b1.double(), groups=8))
    f3 = F.relu(F.conv2d(x.double(), w3.double(), b3.double(), padding=1, groups=8))
    ref = (f1 * f1 + f3 * f3).permute(0, 2, 3, 1).reshape(G * 64, 1024)
    out = torch.zeros((G * 64, 1024), dtype=dtype, device=DEV)
    ops.gemm(xt, HeadEngine._tap_major(w1).contiguous().to(DEV, dtype), out, M=G * 64, N=1024, K=128, block_n=128,
             view="plain", lda=1024, group_c=128, flags=L.EPI_BIAS | L.EPI_RELU | L.EPI_SQUARE, bias=b1.to(DEV))
    ops.gemm(xt, HeadEngine._tap_major(w3).contiguous().to(DEV, dtype), out, M=G * 64, N=1024, K=128, block_n=128,
             view="map", map_args=(1024, 8, 8, 1, G), taps=9, group_c=128,
             flags=L.EPI_BIAS | L.EPI_RELU | L.EPI_SQUARE | L.EPI_ACCUM, bias=b3.to(DEV))
    tol = _tol(dtype)
    if dtype == torch.bfloat16:
        tol = dict(rtol=3e-2, atol=5e-2)      # squared outputs, stored twice in bf16
    torch.testing.assert_close(out.float().cpu().double(), ref, **tol)
    # the same SKBlock as ONE dual-accumulator launch (what the engine runs)
    wf = torch.cat([HeadEngine._tap_major(w3), HeadEngine._tap_major(w1)], dim=1).contiguous().to(DEV, dtype)
    out2 = torch.zeros((G * 64, 1024), dtype=dtype, device=DEV)
    ops.gemm(xt, wf, out2, M=G * 64, N=1024, K=128, block_n=128, view="map", map_args=(1024, 8, 8, 1, G), taps=9,
             group_c=128, flags=L.EPI_BIAS | L.EPI_RELU | L.EPI_SQUARE | L.EPI_DUAL, bias=b3.to(DEV), dual=True,
             bias2=b1.to(DEV))
    torch.testing.assert_close(out2.float().cpu().double(), ref, **tol)


def _attn_ref(q, k, v, w_sk, b_sk, mask):
    """fp64 restatement of Modules.py:16-29 + SubLayers.py:22-39,89-92 for [G, 64, 512] inputs."""
    G = k.shape[0]
    qh = q.view(-1, 64, 8, 64).transpose(1, 2)
    kh = k.view(G, 64, 8, 64).transpose(1, 2)
    vh = v.view(G, 64, 8, 64).transpose(1, 2)
    if qh.shape[0] != G:
        qh = qh.repeat_interleave(G // qh.shape[0], dim=0)
    att = (qh / 8.0) @ kh.transpose(2, 3)
    att = att.masked_fill(mask == 0, -1e9).softmax(-1)
    o = att @ vh
    s = o.sum(1).mean(1)
    gate = (s @ w_sk.t() + b_sk).view(G, 8, 64).softmax(1).unsqueeze(2)
    return (o * gate).sum(1)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("mode", ["self_pad", "causal", "cross"])
def test_attn_core(dtype, mode):
    from ait_b200 import ops
    g = torch.Generator().manual_seed(5)
    G, rep = (6, 1) if mode != "cross" else (6, 3)
    q = _prep(torch.randn(G // rep, 64, 512, generator=g), dtype)
    k = _prep(torch.randn(G, 64, 512, generator=g), dtype)
    v = _prep(torch.randn(G, 64, 512, generator=g), dtype)
    w_sk, b_sk = torch.randn(512, 64, generator=g) * 0.3, torch.randn(512, generator=g) * 0.1
    if mode == "causal":
        mask = torch.tril(torch.ones(64, 64))[None, None]
    else:
        mask = (torch.arange(64) < 49).float()[None, None, None, :]
    ref = _attn_ref(q.double(), k.double(), v.double(), w_sk.double(), b_sk.double(), mask)
    kv = torch.cat([k, v], dim=2).contiguous().to(DEV, dtype)            # [G, 64, 1024], like KVc
    out = torch.zeros((G, 64, 64), dtype=dtype, device=DEV)
    ops.attn_core(q.to(DEV, dtype), 512, rep, kv, kv.view(-1)[512:], 1024, w_sk.to(DEV), b_sk.to(DEV), G,
                  1 if mode == "causal" else 0, 64 if mode == "causal" else 49, out)
    tol = dict(rtol=2e-3, atol=2e-3) if dtype == torch.float32 else dict(rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(out.float().cpu().double(), ref, **tol)


@pytest.mark.parametrize("mode", ["self_pad", "causal", "cross"])
@pytest.mark.parametrize("split", [False, True])
def test_attn_core_tcgen05_variant(mode, split, monkeypatch):
    """The opt-in tcgen05 attention kernel (attn_tc.cu: Q K^T and P V as tcgen05.mma with TMEM accumulators, V consumed
    MN-major, couples of pairs per M = 128 tile) against the same fp64 reference, odd pair counts included."""
    from ait_b200 import ops
    monkeypatch.setenv("AITB_ATTN_TC", "1")
    g = torch.Generator().manual_seed(7)
    G, rep = (7, 1) if mode != "cross" else (9, 3)
    bf = lambda x: x.to(torch.bfloat16).float()                                  # noqa: E731
    q = torch.randn(G // rep, 64, 512, generator=g)
    k = torch.randn(G, 64, 512, generator=g)
    v = torch.randn(G, 64, 512, generator=g)
    if not split:
        q, k, v = bf(q), bf(k), bf(v)
    w_sk, b_sk = torch.randn(512, 64, generator=g) * 0.3, torch.randn(512, generator=g) * 0.1
    if mode == "causal":
        mask = torch.tril(torch.ones(64, 64))[None, None]
    else:
        mask = (torch.arange(64) < 49).float()[None, None, None, :]
    ref = _attn_ref(q.double(), k.double(), v.double(), w_sk.double(), b_sk.double(), mask)
    kv = torch.cat([k, v], dim=2).contiguous()
    if split:
        qd, kvd = ops.split_planes(q).to(DEV), ops.split_planes(kv).to(DEV)
        out = torch.zeros((G * 64, 128), dtype=torch.bfloat16, device=DEV)
        ops.attn_core(qd, 512, rep, kvd, kvd.view(-1)[512:], 1024, w_sk.to(DEV), b_sk.to(DEV), G,
                      1 if mode == "causal" else 0, 64 if mode == "causal" else 49, out, split=True)
        got = ops.join_planes(out).cpu().view(G, 64, 64)
        tol = dict(rtol=1e-4, atol=1e-4)
    else:
        qd, kvd = q.to(DEV, torch.bfloat16), kv.to(DEV, torch.bfloat16)
        out = torch.zeros((G, 64, 64), dtype=torch.bfloat16, device=DEV)
        ops.attn_core(qd, 512, rep, kvd, kvd.view(-1)[512:], 1024, w_sk.to(DEV), b_sk.to(DEV), G,
                      1 if mode == "causal" else 0, 64 if mode == "causal" else 49, out)
        got = out.float().cpu()
        tol = dict(rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(got.double(), ref, **tol)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M", [1000, 128 * 6, 4096 + 37])
@pytest.mark.parametrize("res_relu", [True, False])
def test_gemm_bias_residual_relu_two_cta(dtype, M, res_relu):
    """layer-4 conv3 shape (N = 2048, K = 512): bias + row-for-row residual (+ ReLU after it).  bf16 takes the TMA-store fast
    epilogue with the TMA-loaded residual tile (ragged M: TMA zero-fills / clips the rows past M); guard rows stay untouched."""
    from ait_b200 import _lib as L, ops
    N, K = 2048, 512
    g = torch.Generator().manual_seed(M)
    a = _prep(torch.randn(M, K, generator=g), dtype)
    w = _prep(torch.randn(N, K, generator=g) / K ** 0.5, dtype)
    bias = torch.randn(N, generator=g)
    res = _prep(torch.randn(M, N, generator=g), dtype)
    ref = a.double() @ w.double().t() + bias.double() + res.double()
    if res_relu:
        ref = F.relu(ref)
    out = torch.full((M + 2, N), 5.0, dtype=dtype, device=DEV)
    flags = L.EPI_BIAS | L.EPI_RES | (L.EPI_RES_RELU if res_relu else 0)
    ops.gemm(a.to(DEV, dtype), w.to(DEV, dtype), out, M=M, N=N, K=K, block_n=256, flags=flags, bias=bias.to(DEV),
             res=res.to(DEV, dtype), ldr=N)
    assert bool((out[M:].float() == 5.0).all()), "rows past M were written"
    torch.testing.assert_close(out[:M].float().cpu().double(), ref, **_tol(dtype))
