"""GPU parity of the "fp32-class" split configuration (AITB_F32S): every operand is two bf16 planes
(hi + lo = 16 mantissa bits) and every product is three bf16 tensor-core passes.  References are fp64
computed from the UNROUNDED fp32 operands, so the gates measure the full error of the scheme
(operand split 2^-17 + fp32 accumulation), which must sit ~30x below tf32's 2^-11."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
BF = torch.bfloat16
# max |err| relative to the output's scale.  tf32 measures ~3e-4 on the same problems.
GATE = 2e-5


def _err(out, ref):
    return float((out.double() - ref.double()).abs().max() / ref.double().abs().max())


def _sp(x):
    from ait_b200 import ops
    return ops.split_planes(x).to(DEV)


def _jn(x):
    from ait_b200 import ops
    return ops.join_planes(x).cpu()


def test_split_planes_roundtrip():
    from ait_b200 import ops
    x = torch.randn(64, 256, generator=torch.Generator().manual_seed(0)) * 37.0
    y = ops.join_planes(ops.split_planes(x))
    assert float((x - y).abs().max() / x.abs().max()) < 2 ** -16


@pytest.mark.parametrize("M,N,K,bn", [(128, 256, 64, 256), (300, 512, 512, 256), (1000, 1536, 512, 256),
                                      (77, 128, 128, 128), (4096 + 5, 2048, 512, 256), (640, 256, 2048, 256),
                                      (130, 64, 192, 64)])
def test_split_gemm_plain_bias_relu(M, N, K, bn):
    from ait_b200 import _lib as L, ops
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    bias = torch.randn(N, generator=g)
    ref = F.relu(a.double() @ w.double().t() + bias.double())
    out = torch.full((M, 2 * N), float("nan"), dtype=BF, device=DEV)
    ops.gemm(_sp(a), _sp(w), out, M=M, N=N, K=K, block_n=bn, flags=L.EPI_BIAS | L.EPI_RELU, bias=bias.to(DEV),
             split=True)
    assert _err(_jn(out), ref) < GATE


def test_split_gemm_layernorm_residual_pos_rowmap():
    from ait_b200 import _lib as L, ops
    g = torch.Generator().manual_seed(1)
    pairs, K = 11, 1024
    M = pairs * 49
    a = torch.randn(M, K, generator=g)
    w = torch.randn(512, K, generator=g) / K ** 0.5
    bias, pos = torch.randn(512, generator=g), torch.randn(64, 512, generator=g)
    gamma, beta = torch.rand(512, generator=g) + 0.5, torch.randn(512, generator=g)
    y = (a.double() @ w.double().t() + bias.double()).view(pairs, 49, 512) + pos[:49].double()
    ref = F.layer_norm(y, (512,), gamma.double(), beta.double(), eps=1e-6)
    out = torch.zeros((pairs * 64, 1024), dtype=BF, device=DEV)
    ops.gemm(_sp(a), _sp(w), out, M=M, N=512, K=K, block_n=512, flags=L.EPI_BIAS | L.EPI_POS | L.EPI_LN,
             bias=bias.to(DEV), pos=pos.to(DEV), pos_rows=64, gamma=gamma.to(DEV), beta=beta.to(DEV), rows_in=49,
             rows_out=64, split=True)
    o = _jn(out).view(pairs, 64, 512)
    assert _err(o[:, :49], ref) < GATE
    assert torch.all(o[:, 49:] == 0)


def test_split_gemm_residual_broadcast_layernorm():
    from ait_b200 import _lib as L, ops
    g = torch.Generator().manual_seed(2)
    B, P = 3, 5
    M = B * P * 64
    a = torch.randn(M, 64, generator=g)
    w = torch.randn(512, 64, generator=g) / 8
    res = torch.randn(B * 64, 512, generator=g)
    gamma, beta = torch.rand(512, generator=g) + 0.5, torch.randn(512, generator=g)
    y = (a.double() @ w.double().t()).view(B, P, 64, 512) + res.double().view(B, 1, 64, 512)
    ref = F.layer_norm(y, (512,), gamma.double(), beta.double(), eps=1e-6).view(M, 512)
    out = torch.zeros((M, 1024), dtype=BF, device=DEV)
    ops.gemm(_sp(a), _sp(w), out, M=M, N=512, K=64, block_n=512, flags=L.EPI_RES | L.EPI_LN, res=_sp(res), ldr=512,
             res_div=64, res_rep=P, gamma=gamma.to(DEV), beta=beta.to(DEV), split=True)
    assert _err(_jn(out), ref) < GATE


@pytest.mark.parametrize("G", [1, 19])
def test_split_gemm_convs(G):
    """3x3 on a 4x4 map (shifted TMA boxes), stride-2 1x1, residual + relu tail, grouped dual-accumulator SKBlock."""
    from ait_b200 import _lib as L, ops
    from ait_b200.packing import HeadEngine
    g = torch.Generator().manual_seed(G)
    x = torch.randn(G, 512, 4, 4, generator=g)
    w = torch.randn(512, 512, 3, 3, generator=g) / 68.0
    bias = torch.randn(512, generator=g)
    ref = F.relu(F.conv2d(x.double(), w.double(), bias.double(), padding=1)).permute(0, 2, 3, 1).reshape(G * 16, 512)
    out = torch.zeros((G * 16, 1024), dtype=BF, device=DEV)
    ops.gemm(_sp(x.permute(0, 2, 3, 1).contiguous()), _sp(HeadEngine._tap_major(w).contiguous()), out, M=G * 16,
             N=512, K=512, block_n=256, view="map", map_args=(512, 4, 4, 1, G), taps=9,
             flags=L.EPI_BIAS | L.EPI_RELU, bias=bias.to(DEV), split=True)
    assert _err(_jn(out), ref) < GATE

    x8 = torch.randn(G, 1024, 8, 8, generator=g)
    xt = _sp(x8.permute(0, 2, 3, 1).contiguous())
    w1 = torch.randn(2048, 1024, 1, 1, generator=g) / 32.0
    res = torch.randn(G * 16, 2048, generator=g)
    ref = F.relu(F.conv2d(x8.double(), w1.double(), stride=2).permute(0, 2, 3, 1).reshape(G * 16, 2048) + res.double())
    out = torch.zeros((G * 16, 4096), dtype=BF, device=DEV)
    ops.gemm(xt, _sp(w1.flatten(1).contiguous()), out, M=G * 16, N=2048, K=1024, block_n=256, view="map",
             map_args=(1024, 8, 4, 2, G), flags=L.EPI_RES | L.EPI_RES_RELU, res=_sp(res), ldr=2048, split=True)
    assert _err(_jn(out), ref) < GATE

    k1 = torch.randn(1024, 128, 1, 1, generator=g) / 11.0
    k3 = torch.randn(1024, 128, 3, 3, generator=g) / 34.0
    b1, b3 = torch.randn(1024, generator=g) * 0.1, torch.randn(1024, generator=g) * 0.1
    f1 = F.relu(F.conv2d(x8.double(), k1.double(), b1.double(), groups=8))
    f3 = F.relu(F.conv2d(x8.double(), k3.double(), b3.double(), padding=1, groups=8))
    ref = (f1 * f1 + f3 * f3).permute(0, 2, 3, 1).reshape(G * 64, 1024)
    wf = torch.cat([HeadEngine._tap_major(k3), HeadEngine._tap_major(k1)], dim=1).contiguous()
    out = torch.zeros((G * 64, 2048), dtype=BF, device=DEV)
    ops.gemm(xt, _sp(wf), out, M=G * 64, N=1024, K=128, block_n=128, view="map", map_args=(1024, 8, 8, 1, G), taps=9,
             group_c=128, flags=L.EPI_BIAS | L.EPI_RELU | L.EPI_SQUARE | L.EPI_DUAL, bias=b3.to(DEV), dual=True,
             bias2=b1.to(DEV), split=True)
    assert _err(_jn(out), ref) < 2 * GATE        # squared outputs double the relative error


@pytest.mark.parametrize("mode", ["self_pad", "causal", "cross"])
def test_split_attn_core(mode):
    from ait_b200 import ops
    from test_gpu_gemm_attn import _attn_ref
    g = torch.Generator().manual_seed(5)
    G, rep = (6, 1) if mode != "cross" else (6, 3)
    q = torch.randn(G // rep, 64, 512, generator=g)
    k = torch.randn(G, 64, 512, generator=g)
    v = torch.randn(G, 64, 512, generator=g)
    w_sk, b_sk = torch.randn(512, 64, generator=g) * 0.3, torch.randn(512, generator=g) * 0.1
    if mode == "causal":
        mask = torch.tril(torch.ones(64, 64))[None, None]
    else:
        mask = (torch.arange(64) < 49).float()[None, None, None, :]
    ref = _attn_ref(q.double(), k.double(), v.double(), w_sk.double(), b_sk.double(), mask)
    kv = _sp(torch.cat([k, v], dim=2).contiguous())                      # [G, 64, hi 1024 | lo 1024], like KVc
    out = torch.zeros((G, 64, 128), dtype=BF, device=DEV)
    ops.attn_core(_sp(q), 512, rep, kv, kv.view(-1)[512:], 1024, w_sk.to(DEV), b_sk.to(DEV), G,
                  1 if mode == "causal" else 0, 64 if mode == "causal" else 49, out, split=True)
    assert _err(_jn(out), ref) < 1e-4            # tf32/fp16 path: ~2e-3


def test_split_transposes_and_pool_heads():
    from ait_b200 import ops
    g = torch.Generator().manual_seed(9)
    x = torch.randn(3, 1024, 64, generator=g)
    t = ops.transpose_cs(x.to(DEV), True, split_dst=True)
    assert t.shape == (3, 64, 2048) and t.dtype == BF
    assert _err(_jn(t), x.transpose(1, 2)) < 2 ** -16
    back = ops.transpose_cs(t, False, split_src=True)
    assert back.dtype == torch.float32 and _err(back.cpu(), x) < 2 ** -16
    top = torch.randn(5, 16, 2048, generator=g)
    feat, _, _ = ops.pool_heads(_sp(top), 1, split=True)
    assert _err(feat.cpu(), top.mean(1)) < 1e-5


@pytest.mark.parametrize("split", [True, False])
@pytest.mark.parametrize("case", ["plain", "broadcast", "compact"])
def test_fc_ln_streaming_kernel(split, case):
    """fc (64 -> 512) + residual + LayerNorm as the streaming kernel (fc_ln.cu): plain rows, the cross-attention
    residual broadcast (one residual block per unit, P pairs), and the encoder's 64 -> 49 row compaction with the
    residual indexed by the 64-row layout; ragged M (not a multiple of the 64-row CTA block)."""
    from ait_b200 import ops
    g = torch.Generator().manual_seed(5)
    B, P = 3, 5
    pairs = B * P
    M = pairs * 64
    a = torch.randn(M, 64, generator=g)
    w = torch.randn(512, 64, generator=g) / 8
    gamma, beta = torch.rand(512, generator=g) + 0.5, torch.randn(512, generator=g)
    kw = {}
    if case == "broadcast":
        res = torch.randn(B * 64, 512, generator=g)
        y = (a.double() @ w.double().t()).view(B, P, 64, 512) + res.double().view(B, 1, 64, 512)
        kw = dict(res_div=64, res_rep=P)
        rows = M
    elif case == "compact":
        res = torch.randn(M, 512, generator=g)
        y = ((a.double() @ w.double().t()) + res.double()).view(pairs, 64, 512)[:, :49]
        kw = dict(rows_in=64, rows_out=49, res_row_m=True, res_div=64, res_rep=1)
        rows = pairs * 49
    else:
        M = M - 40          # ragged tail
        a, res = a[:M], torch.randn(M, 512, generator=g)
        y = a.double() @ w.double().t() + res.double()
        kw = dict(res_div=64, res_rep=1)
        rows = M
    ref = F.layer_norm(y, (512,), gamma.double(), beta.double(), eps=1e-6).reshape(rows, 512)
    conv = _sp if split else (lambda x: x.to(BF).to(DEV))
    out = torch.zeros((rows + 3, 1024 if split else 512), dtype=BF, device=DEV)
    ops.fc_ln(conv(a), conv(w), conv(res), gamma.to(DEV), beta.to(DEV), out, M=M, split=split, **kw)
    o = _jn(out) if split else out.float().cpu()
    assert torch.all(o[rows:] == 0), "rows past the output were written"
    if split:
        assert _err(o[:rows], ref) < GATE
    else:   # bf16 storage: compare against the same computation on bf16-rounded operands, bf16 output rounding
        ab, wb, rb = a.to(BF).double(), w.to(BF).double(), res.to(BF).double()
        if case == "broadcast":
            yb = (ab @ wb.t()).view(B, P, 64, 512) + rb.view(B, 1, 64, 512)
        elif case == "compact":
            yb = ((ab @ wb.t()) + rb).view(pairs, 64, 512)[:, :49]
        else:
            yb = ab @ wb.t() + rb
        refb = F.layer_norm(yb, (512,), gamma.double(), beta.double(), eps=1e-6).reshape(rows, 512)
        assert _err(o[:rows], refb) < 6e-3


# ---------------------------------------------------------------------------------------------
# precision plan: one tensor-core pass on fp16 hi planes (aitb_gemm_desc.passes = 1, in_f16 / out_f16 / res_f16)
# ---------------------------------------------------------------------------------------------
def _sp16(x):
    from ait_b200 import ops
    return ops.split_planes(x, f16=True).to(DEV)


def _jn16(x):
    from ait_b200 import ops
    return ops.join_planes(x, f16=True).cpu()


def test_fp16_planes_roundtrip_and_saturation():
    from ait_b200 import ops
    x = torch.randn(64, 256, generator=torch.Generator().manual_seed(0)) * 37.0
    y = ops.join_planes(ops.split_planes(x, f16=True), f16=True)
    assert float((x - y).abs().max() / x.abs().max()) < 2 ** -20          # 11 + 11 bits
    big = torch.tensor([[1e6, -3e5, 70000.0, 1.0]])
    z = ops.join_planes(ops.split_planes(big, f16=True), f16=True)
    assert torch.isfinite(z).all() and float(z[0, 0]) == 65504.0 and float(z[0, 3]) == 1.0


@pytest.mark.parametrize("M,N,K", [(300, 512, 512), (1000, 1536, 512), (4096 + 5, 2048, 512), (640, 256, 2048), (128, 256, 64)])
def test_onepass_fp16_gemm_plain_bias_relu(M, N, K):
    """2-CTA kernel, one pass: the result equals the exact product of the fp16-ROUNDED operands (the hi planes) to fp32
    accumulation error, and the full-precision product to the 11-bit operand rounding (~tf32 class); the output is written
    as fp16 planes.  M = 128 (one m-tile) takes the three-pass single-CTA kernel on the same fp16 planes."""
    from ait_b200 import _lib as L, ops
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    bias = torch.randn(N, generator=g)
    ref_full = F.relu(a.double() @ w.double().t() + bias.double())
    ref_hi = F.relu(a.half().double() @ w.half().double().t() + bias.double())
    out = torch.full((M, 2 * N), float("nan"), dtype=BF, device=DEV)
    ops.gemm(_sp16(a), _sp16(w), out, M=M, N=N, K=K, block_n=256, flags=L.EPI_BIAS | L.EPI_RELU, bias=bias.to(DEV),
             split=True, passes=1, in_f16=True, out_f16=True)
    got = _jn16(out)
    if M >= 256:
        assert _err(got, ref_hi) < 4e-6
        assert 2e-5 < _err(got, ref_full) < 2e-3      # really one pass: the operand rounding is visible
    else:
        assert _err(got, ref_full) < 4e-6             # three passes on fp16 planes: 22-bit operands


def test_onepass_fp16_gemm_layernorm_residual():
    """cluster-LayerNorm kernel, one pass: FFN w_2 shape with an fp16-plane residual and fp16-plane output."""
    from ait_b200 import _lib as L, ops
    g = torch.Generator().manual_seed(5)
    M, K = 700, 2048
    a = torch.relu(torch.randn(M, K, generator=g))
    w = torch.randn(512, K, generator=g) / K ** 0.5
    bias, res = torch.randn(512, generator=g), torch.randn(M, 512, generator=g)
    gamma, beta = 1 + 0.1 * torch.randn(512, generator=g), 0.1 * torch.randn(512, generator=g)
    x = a.half().double() @ w.half().double().t() + bias.double() + res.double()
    ref = F.layer_norm(x, (512,), gamma.double(), beta.double(), eps=1e-6)
    out = torch.full((M, 1024), float("nan"), dtype=BF, device=DEV)
    ops.gemm(_sp16(a), _sp16(w), out, M=M, N=512, K=K, block_n=512, flags=L.EPI_BIAS | L.EPI_RES | L.EPI_LN,
             bias=bias.to(DEV), res=_sp16(res), ldr=512, gamma=gamma.to(DEV), beta=beta.to(DEV), split=True,
             passes=1, in_f16=True, out_f16=True, res_f16=True)
    assert _err(_jn16(out), ref) < 1e-5
    # mixed formats: bf16-plane residual, bf16-plane output, fp16 operands
    out2 = torch.full((M, 1024), float("nan"), dtype=BF, device=DEV)
    ops.gemm(_sp16(a), _sp16(w), out2, M=M, N=512, K=K, block_n=512, flags=L.EPI_BIAS | L.EPI_RES | L.EPI_LN,
             bias=bias.to(DEV), res=_sp(res), ldr=512, gamma=gamma.to(DEV), beta=beta.to(DEV), split=True,
             passes=1, in_f16=True)
    assert _err(_jn(out2), ref) < 3e-5


@pytest.mark.parametrize("flags_name", ["none", "bias_relu", "bias_relu_hi_only"])
def test_onepass_fast_epilogue_matches_general_path(flags_name, monkeypatch):
    """The TMA-store fast epilogue of the one-pass 2-CTA kernel (hi | lo planes, fp16) against the exact product of the hi
    planes, ragged M (TMA clips the last rows) and, with HI_ONLY, an untouched lo plane."""
    from ait_b200 import _lib as L, ops
    M, N, K = 128 * 7 + 45, 1536, 512
    g = torch.Generator().manual_seed(9)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    bias = torch.randn(N, generator=g)
    flags = 0 if flags_name == "none" else (L.EPI_BIAS | L.EPI_RELU)
    if flags_name.endswith("hi_only"):
        flags |= L.EPI_HI_ONLY
    x = a.half().double() @ w.half().double().t()
    ref = F.relu(x + bias.double()) if flags else x
    out = torch.full((M + 3, 2 * N), 7.0, dtype=BF, device=DEV)          # 3 guard rows past M
    ops.gemm(_sp16(a), _sp16(w), out, M=M, N=N, K=K, block_n=256, flags=flags, bias=bias.to(DEV) if flags else None,
             split=True, passes=1, in_f16=True, out_f16=True)
    assert bool((out[M:].float() == 7.0).all()), "rows past M were written"
    if flags_name.endswith("hi_only"):
        assert bool((out[:M, N:].float() == 7.0).all()), "HI_ONLY wrote the lo plane"
        hi = out[:M, :N].contiguous().view(torch.float16).float().cpu()
        assert _err(hi, ref) < 1e-3                                       # 11-bit plane alone
    else:
        assert _err(_jn16(out[:M]), ref) < 4e-6
