"""GPU parity of the training path (BASELINE config 4): backward building blocks and the AIT
forward + backward against the CPU oracle's autograd (fp64) and the reference's own gradients
(tests/golden/ait_grad.pt, generated from the unmodified reference modules)."""
import re

import numpy as np
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


@pytest.mark.parametrize("M,N,K", [(64, 128, 64), (1000, 512, 64), (4096 + 37, 1536, 512), (777, 512, 1024),
                                   (20000, 2048, 512), (333, 512, 2048), (31, 128, 256), (6, 512, 64)])
def test_wgrad_mn_major_tcgen05_bf16(M, N, K):
    """bf16 training configuration: dW = dY^T X from bf16 row-major activations (MN-major bf16 tiles, 128-byte swizzle,
    kind::f16 MMAs of K = 16), fp32 accumulation: exact products of the bf16 operands up to fp32 summation order."""
    from ait_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    dy = torch.randn(M, N, generator=g).to(torch.bfloat16)
    x = torch.randn(M, K, generator=g).to(torch.bfloat16)
    ref = dy.double().t() @ x.double()
    out = ops.wgrad(dy.to(DEV), x.to(DEV))
    assert out.dtype == torch.float32
    err = float((out.cpu().double() - ref).abs().max() / ref.abs().max())
    assert err < 2e-5, err
    wide = torch.randn(M, N + 256, generator=g).to(torch.bfloat16).to(DEV)
    base = torch.ones(N, K, device=DEV)
    out2 = ops.wgrad(wide[:, 128:], x.to(DEV), dw=base.clone(), N=N)
    ref2 = 1.0 + wide[:, 128:128 + N].cpu().double().t() @ x.double()
    assert float((out2.cpu().double() - ref2).abs().max() / ref2.abs().max()) < 2e-5


def test_ln_backward_from_saved_output():
    from ait_b200 import ops
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(3)
    pairs = 5
    x = torch.randn(pairs * 64, 512, generator=g, dtype=torch.float64, requires_grad=True)
    gamma = (torch.rand(512, generator=g, dtype=torch.float64) + 0.5).requires_grad_()
    beta = torch.randn(512, generator=g, dtype=torch.float64).requires_grad_()
    y = F.layer_norm(x, (512,), gamma, beta, eps=1e-6)
    gy = torch.randn(pairs * 64, 512, generator=g, dtype=torch.float64)
    y.backward(gy)
    rstd = 1.0 / torch.sqrt(x.detach().var(dim=1, unbiased=False) + 1e-6)
    dx, dgamma, dbeta = ops.ln_bwd(gy.float().to(DEV), y.detach().float().to(DEV), gamma.detach().float().to(DEV),
                                   beta.detach().float().to(DEV), rstd.float().to(DEV))
    # dx is rounded to tf32 where it is produced (it feeds tf32 MMAs next): 2^-11 relative
    assert float((dx.cpu().double() - x.grad).abs().max() / x.grad.abs().max()) < 6e-4
    assert float((dgamma.cpu().double() - gamma.grad).abs().max() / gamma.grad.abs().max()) < 1e-4
    assert float((dbeta.cpu().double() - beta.grad).abs().max() / beta.grad.abs().max()) < 1e-4
    # encoder un-padding: only the first 49 rows of every 64-row group produce dx (compacted), all rows feed dgamma / dbeta
    dx49, dgamma2, _ = ops.ln_bwd(gy.float().to(DEV), y.detach().float().to(DEV), gamma.detach().float().to(DEV),
                                  beta.detach().float().to(DEV), rstd.float().to(DEV), grp=64, valid=49)
    ref49 = x.grad.view(pairs, 64, 512)[:, :49].reshape(-1, 512)
    assert dx49.shape == (pairs * 49, 512)
    assert float((dx49.cpu().double() - ref49).abs().max() / ref49.abs().max()) < 6e-4
    assert float((dgamma2.cpu().double() - gamma.grad).abs().max() / gamma.grad.abs().max()) < 1e-4


def test_colsum_and_bsum():
    from ait_b200 import ops
    g = torch.Generator().manual_seed(4)
    x = torch.randn(3000, 2048, generator=g)
    assert torch.allclose(ops.colsum(x.to(DEV)).cpu().double(), x.double().sum(0), rtol=1e-4, atol=1e-3)
    y = torch.randn(3, 7, 64 * 512, generator=g)
    assert torch.allclose(ops.bsum(y.to(DEV)).cpu().double(), y.double().sum(1), rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("mode", ["self_pad", "causal", "cross"])
def test_attention_backward(mode):
    """selective-head attention backward against fp64 autograd of the reference formulation."""
    from ait_b200 import ops
    g = torch.Generator().manual_seed(6)
    G, rep = (5, 1) if mode != "cross" else (6, 3)
    q = _tf32(torch.randn(G // rep, 64, 512, generator=g)).double().requires_grad_()
    k = _tf32(torch.randn(G, 64, 512, generator=g)).double().requires_grad_()
    v = _tf32(torch.randn(G, 64, 512, generator=g)).double().requires_grad_()
    w_sk = (torch.randn(512, 64, generator=g) * 0.3).double().requires_grad_()
    b_sk = (torch.randn(512, generator=g) * 0.1).double().requires_grad_()
    if mode == "causal":
        mask = torch.tril(torch.ones(64, 64))[None, None]
    else:
        mask = (torch.arange(64) < 49).float()[None, None, None, :]
    qh = q.view(-1, 64, 8, 64).transpose(1, 2)
    if rep > 1:
        qh = qh.repeat_interleave(rep, dim=0)
    kh = k.view(G, 64, 8, 64).transpose(1, 2)
    vh = v.view(G, 64, 8, 64).transpose(1, 2)
    att = ((qh / 8.0) @ kh.transpose(2, 3)).masked_fill(mask == 0, -1e9).softmax(-1)
    o = att @ vh
    s = o.sum(1).mean(1)
    gate = (s @ w_sk.t() + b_sk).view(G, 8, 64).softmax(1).unsqueeze(2)
    out = (o * gate).sum(1)
    dout = _tf32(torch.randn(G, 64, 64, generator=g)).double()
    out.backward(dout)
    kv = torch.cat([k.detach(), v.detach()], dim=2).float().contiguous().to(DEV)
    dq, dk, dv, dz, sv = ops.attn_bwd(q.detach().float().to(DEV), 512, rep, kv, kv.view(-1)[512:], 1024,
                                      w_sk.detach().float().to(DEV), b_sk.detach().float().to(DEV),
                                      dout.float().to(DEV), G, 1 if mode == "causal" else 0,
                                      64 if mode == "causal" else 49)
    dq = dq.view(G // rep, rep, 64, 512).sum(1).cpu().double()

    def rel(a, b):
        return float((a - b).abs().max() / b.abs().max())

    assert rel(dq, q.grad) < 3e-3
    assert rel(dk.view(G, 64, 512).cpu().double(), k.grad) < 3e-3
    assert rel(dv.view(G, 64, 512).cpu().double(), v.grad) < 3e-3
    # d(W_sk) = dz^T s, d(b_sk) = colsum(dz)
    assert rel(dz.cpu().double().t() @ sv.cpu().double(), w_sk.grad) < 3e-3
    assert rel(dz.cpu().double().sum(0), b_sk.grad) < 3e-3


def _oracle_grads(m, x_props, x_query, gout):
    from oracle import head_oracle
    sd = {k: v.detach().double().clone().requires_grad_(v.is_floating_point() and "pos_table" not in k)
          for k, v in m.state_dict().items()}
    xp, xq = x_props.double().requires_grad_(), x_query.double().requires_grad_()
    ref = head_oracle.ait_forward(sd, xp, xq, dtype=torch.float64)
    ref.backward(gout.double())
    return ref.detach(), xp.grad, xq.grad, {k: v.grad for k, v in sd.items() if v.grad is not None}


def _l2rel(a, b):
    a = a.detach().cpu().double()
    return float((a - b).norm() / b.norm())


@pytest.mark.parametrize("B,P,smooth", [(1, 2, True), (2, 3, True), (2, 3, False)])
def test_ait_training_step_matches_oracle_autograd(B, P, smooth):
    """Transformer forward + backward (config 4, dropout 0): output, both input gradients and all 46 parameter
    gradients against fp64 autograd of the CPU oracle (oracle/head_oracle.py, pinned to the reference).

    smooth: the FFN biases are shifted so that every ReLU is active -- the loss is then smooth in all weights
    and a tf32 forward/backward must agree with fp64 to tf32 accuracy (gate 6e-3 relative L2; measured <= 3.4e-3).
    With the stock init a tf32 forward flips the ReLU mask of the ~0.1 % of hidden units whose pre-activation is
    within tf32 error of zero, which alone is a sqrt(1e-3) ~ 2-3 % relative-L2 difference of everything upstream
    of the FFNs (measured 1.8e-2): gate 4e-2 there."""
    from ait_b200.system.Models import Transformer
    torch.manual_seed(0)
    m = Transformer(n_layers=1, dropout=0.0, n_position=64, attn_dropout=0.0).train()
    if smooth:
        with torch.no_grad():
            m.encoder.layer_stack[0].pos_ffn.w_1.bias += 5.0
            m.decoder.layer_stack[0].pos_ffn.w_1.bias += 5.0
    g = torch.Generator().manual_seed(11 + B)
    x_props = torch.randn(B * P, 1024, 7, 7, generator=g).relu()
    x_query = torch.randn(B, 1024, 8, 8, generator=g).relu()
    gout = torch.randn(B * P, 1024, 8, 8, generator=g)
    ref, gp_ref, gq_ref, pg_ref = _oracle_grads(m, x_props, x_query, gout)
    m = m.to(DEV)
    xp2, xq2 = x_props.to(DEV).requires_grad_(), x_query.to(DEV).requires_grad_()
    out = m(xp2, xq2)
    out.backward(gout.to(DEV))
    torch.cuda.synchronize()
    gate = 6e-3 if smooth else 4e-2
    assert _l2rel(out, ref) < 2e-3
    assert _l2rel(xp2.grad, gp_ref) < gate, "grad x_props"
    assert _l2rel(xq2.grad, gq_ref) < gate, "grad x_query"
    errs = {}
    for name, p in m.named_parameters():
        assert p.grad is not None, name
        errs[name] = _l2rel(p.grad, pg_ref[name])
    bad = {k: v for k, v in errs.items() if v > gate}
    assert len(errs) == 46 and not bad, bad
    # a second backward through a fresh forward accumulates into .grad like torch autograd does
    out2 = m(xp2, xq2)
    out2.backward(gout.to(DEV))
    name = "decoder.layer_stack.0.pos_ffn.w_2.weight"
    assert _l2rel(dict(m.named_parameters())[name].grad, 2 * pg_ref[name]) < gate


@pytest.mark.parametrize("p,p_attn,B,P", [(0.1, 0.1, 2, 3), (0.0, 0.1, 1, 2), (0.25, 0.0, 2, 2)])
def test_ait_training_step_with_dropout_matches_oracle_autograd(p, p_attn, B, P):
    """.train() WITH dropout (the reference's default training mode: dropout = 0.1 at the seven nn.Dropout sites and the
    hard-wired 0.1 on the attention probabilities): the device regenerates its masks from (seed, site, element) in forward
    and backward and never stores them; `HeadEngine.dropout_masks` materialises the same multipliers, which are injected
    into the fp64 oracle (whose ten dropout sites are pinned to the reference's nn.Dropout instances by
    tests/test_oracle_pins.py::test_oracle_dropout_sites_match_reference_golden).  Output, both input gradients and all
    46 parameter gradients must then agree to tf32 accuracy (smooth FFN: every ReLU active, see the test above)."""
    from ait_b200 import packing
    from ait_b200.system.Models import Transformer
    from oracle import head_oracle
    torch.manual_seed(0)
    m = Transformer(n_layers=1, dropout=p, n_position=64, attn_dropout=p_attn).train()
    with torch.no_grad():
        m.encoder.layer_stack[0].pos_ffn.w_1.bias += 5.0
        m.decoder.layer_stack[0].pos_ffn.w_1.bias += 5.0
    g = torch.Generator().manual_seed(31 + B)
    x_props = torch.randn(B * P, 1024, 7, 7, generator=g).relu()
    x_query = torch.randn(B, 1024, 8, 8, generator=g).relu()
    gout = torch.randn(B * P, 1024, 8, 8, generator=g)
    sd = {k: v.detach().double().clone().requires_grad_(v.is_floating_point() and "pos_table" not in k)
          for k, v in m.state_dict().items()}
    m = m.to(DEV)
    xp2, xq2 = x_props.to(DEV).requires_grad_(), x_query.to(DEV).requires_grad_()
    torch.manual_seed(123)
    out = m(xp2, xq2)
    out.backward(gout.to(DEV))
    torch.cuda.synchronize()
    seed = m.last_dropout_seed
    assert seed != 0
    eng = packing.HeadEngine(transformer=m, dtype="tf32")
    eng.set_train_dropout(p, p_attn, seed)
    masks = {k: v.cpu() for k, v in eng.dropout_masks(B, P, torch.device(DEV)).items()}
    for name, mk in masks.items():          # Bernoulli(1 - q) multipliers 0 | 1 / (1 - q)
        q = p_attn if name.endswith("_attn") else p
        if q == 0.0:
            assert bool((mk == 1.0).all()), name
            continue
        keep = mk > 0
        assert torch.allclose(mk[keep], torch.full_like(mk[keep], 1.0 / (1.0 - q)), rtol=1e-6), name
        frac = float(keep.float().mean())
        assert abs(frac - (1.0 - q)) < 4.0 * (q * (1 - q) / mk.numel()) ** 0.5 + 1e-3, (name, frac)
    if p > 0.0:
        assert not torch.equal(masks["enc_slf_fc"], masks["enc_ffn"])       # one key per site
    if p_attn > 0.0:
        assert not torch.equal(masks["enc_slf_attn"], masks["dec_enc_attn"])
    xp, xq = x_props.double().requires_grad_(), x_query.double().requires_grad_()
    ref = head_oracle.ait_forward(sd, xp, xq, dtype=torch.float64, drop=masks)
    ref.backward(gout.double())
    with torch.no_grad():
        plain = head_oracle.ait_forward(sd, xp, xq, dtype=torch.float64)
    # the masks move the result beyond the 2e-3 output gate (attention-probability dropout alone: 5.6e-3; with the row sites: > 5e-2)
    assert _l2rel(out, plain) > (5e-2 if p > 0.0 else 4e-3)
    gate = 6e-3
    assert _l2rel(out, ref.detach()) < 2e-3
    assert _l2rel(xp2.grad, xp.grad) < gate, "grad x_props"
    assert _l2rel(xq2.grad, xq.grad) < gate, "grad x_query"
    errs = {name: _l2rel(prm.grad, sd[name].grad) for name, prm in m.named_parameters()}
    bad = {k: v for k, v in errs.items() if v > gate}
    assert len(errs) == 46 and not bad, bad
    # same seed -> same masks -> same step; a new seed -> a different one
    m.zero_grad()
    torch.manual_seed(123)
    out_b = m(xp2, xq2)
    assert m.last_dropout_seed == seed and torch.equal(out_b, out)
    out_c = m(xp2, xq2)
    assert m.last_dropout_seed != seed and not torch.equal(out_c, out)


@pytest.mark.parametrize("B,P,p,p_attn", [(2, 3, 0.0, 0.0), (2, 2, 0.1, 0.1)])
def test_ait_training_step_bf16_matches_oracle_autograd(B, P, p, p_attn):
    """BASELINE configs[3] "fp32/bf16": the bf16 training configuration (Transformer(compute_dtype=torch.bfloat16).train():
    bf16 storage of every activation and gradient, bf16 tcgen05 GEMMs for the forward, dgrad and wgrad products, fp32
    accumulation / LayerNorm statistics / softmax / parameter gradients) against fp64 autograd of the oracle, without and
    with dropout (the device's masks injected into the oracle).  Smooth FFN (every ReLU active).  bf16 carries 8
    significand bits: the stated tolerance is 4e-2 relative L2 per tensor (measured worst 2.5e-2 without / 1.4e-2 with
    dropout; the tf32 configuration: 6e-3)."""
    from ait_b200 import packing
    from ait_b200.system.Models import Transformer
    from oracle import head_oracle
    torch.manual_seed(0)
    m = Transformer(n_layers=1, dropout=p, n_position=64, attn_dropout=p_attn, compute_dtype=torch.bfloat16).train()
    with torch.no_grad():
        m.encoder.layer_stack[0].pos_ffn.w_1.bias += 5.0
        m.decoder.layer_stack[0].pos_ffn.w_1.bias += 5.0
    g = torch.Generator().manual_seed(51 + B)
    x_props = torch.randn(B * P, 1024, 7, 7, generator=g).relu()
    x_query = torch.randn(B, 1024, 8, 8, generator=g).relu()
    gout = torch.randn(B * P, 1024, 8, 8, generator=g)
    sd = {k: v.detach().double().clone().requires_grad_(v.is_floating_point() and "pos_table" not in k)
          for k, v in m.state_dict().items()}
    m = m.to(DEV)
    xp2, xq2 = x_props.to(DEV).requires_grad_(), x_query.to(DEV).requires_grad_()
    out = m(xp2, xq2)
    assert out.dtype == torch.float32
    out.backward(gout.to(DEV))
    torch.cuda.synchronize()
    masks = None
    if p > 0 or p_attn > 0:
        eng = packing.HeadEngine(transformer=m, dtype="bf16")
        eng.set_train_dropout(p, p_attn, m.last_dropout_seed)
        masks = {k: v.cpu() for k, v in eng.dropout_masks(B, P, torch.device(DEV)).items()}
    xp, xq = x_props.double().requires_grad_(), x_query.double().requires_grad_()
    ref = head_oracle.ait_forward(sd, xp, xq, dtype=torch.float64, drop=masks)
    ref.backward(gout.double())
    gate = 4e-2
    errs = {"out": _l2rel(out, ref.detach()), "grad x_props": _l2rel(xp2.grad, xp.grad), "grad x_query": _l2rel(xq2.grad, xq.grad)}
    for name, prm in m.named_parameters():
        assert prm.grad is not None and prm.grad.dtype == torch.float32, name
        errs[name] = _l2rel(prm.grad, sd[name].grad)
    print("bf16 training step: worst relative L2 %.3e (%s)" % (max(errs.values()), max(errs, key=errs.get)))
    bad = {k: v for k, v in errs.items() if not v < gate}
    assert len(errs) == 49 and not bad, bad


def test_ait_training_step_matches_reference_golden_gradients():
    """tests/golden/ait_grad.pt: gradients produced by the UNMODIFIED reference Transformer (CPU fp32 autograd,
    dropout 0.0) with the weights of synth.make_head(seed=0) -- see tests/golden/make_golden_grad.py."""
    from conftest import load_golden
    from ait_b200 import synth
    gold = load_golden("ait_grad.pt")
    g = torch.Generator().manual_seed(13)
    xp = torch.rand(2, 1024, 7, 7, generator=g)
    xq = torch.rand(1, 1024, 8, 8, generator=g)
    gout = torch.randn(2, 1024, 8, 8, generator=g)
    head = synth.make_head(seed=0, calibrated=True, randomize_bn=True)
    m = head.transformer
    from ait_b200.system.Models import set_dropout
    set_dropout(m, 0.0, 0.0)
    m = m.to(DEV).train()
    xp2, xq2 = xp.to(DEV).requires_grad_(), xq.to(DEV).requires_grad_()
    out = m(xp2, xq2)
    out.backward(gout.to(DEV))
    torch.cuda.synchronize()
    assert _l2rel(out[:, ::8], gold["out_s"].double()) < 2e-3
    assert _l2rel(xp2.grad[:, ::4], gold["grad_props_s"].double()) < 4e-2
    assert _l2rel(xq2.grad[:, ::4], gold["grad_query_s"].double()) < 4e-2
    params = dict(m.named_parameters())
    assert set(gold["params"]) == set(params)
    for name, ref in gold["params"].items():
        gr = params[name].grad.reshape(-1)
        sample = gr[::ref["stride"]][:ref["sample"].numel()]
        assert _l2rel(sample, ref["sample"].double()) < 6e-2, name
        assert abs(float(gr.double().norm()) / ref["norm"] - 1.0) < 4e-2, name


@pytest.mark.parametrize("G", [19, 48])
def test_head_to_tail_training_matches_oracle_autograd(G):
    """`_head_to_tail` (layer4 with frozen, randomised BatchNorm + 4x4 mean) forward and backward on the device
    (tf32 tensor-core math): output, input gradient and the gradients of all ten convolution weights; ragged row
    counts (G*16 not a multiple of the 128-row tiles).  Two references: (1) fp64 autograd over the oracle restatement
    -- gate 6e-2, because a tf32 forward flips the ReLU mask of the units whose pre-activation is within tf32 error
    of zero (nine ReLUs deep; measured 0.7 % at the last block growing to 4.7 % at the input); (2) the same fp64 graph
    with ReLU(x) = x * mask and the masks taken from the device forward -- gate 2e-3 (measured 4-8e-4): the
    arithmetic itself."""
    import torch.nn.functional as F
    from ait_b200 import synth, top_train
    from ait_b200.top_train import head_to_tail_train
    from oracle import head_oracle
    head = synth.make_head(seed=0, calibrated=True, randomize_bn=True)
    top = head.RCNN_top
    sd = {k: v.clone() for k, v in top.state_dict().items()}
    g = torch.Generator().manual_seed(33)
    x = torch.randn(G, 1024, 8, 8, generator=g).relu()
    gy = torch.randn(G, 2048, generator=g)
    names = [k for k in sd if k.endswith("weight") and sd[k].dim() == 4]
    assert len(names) == 10
    ref_w = {k: sd[k].double().requires_grad_() for k in names}
    xr = x.double().requires_grad_()
    sd64 = {k: (ref_w[k] if k in ref_w else v.double()) for k, v in sd.items()}
    ref = head_oracle.head_to_tail(sd64, xr, dtype=torch.float64)
    ref.backward(gy.double())
    # device
    top = top.to(DEV)
    for p in top.parameters():
        p.requires_grad_(p.dim() == 4)
    xd = x.to(DEV).requires_grad_()
    out = head_to_tail_train(top, xd)
    out.backward(gy.to(DEV))
    got = dict(top.named_parameters())

    def rel(a, b):
        return float((a.double().cpu() - b).norm() / b.norm())

    assert rel(out.detach(), ref.detach()) < 2e-3
    assert rel(xd.grad, xr.grad) < 6e-2
    for k in names:
        assert rel(got[k].grad, ref_w[k].grad) < 6e-2, k
    # (2) the reference with the device's masks
    masks = [[(t > 0).double().cpu() for t in blk] for blk in top_train._last_saved_for_tests]

    def tok2map(m):  # [G*16, C] token-major -> [G, C, 4, 4]
        return m.view(G, 4, 4, -1).permute(0, 3, 1, 2)

    def bn(t, pre):
        w, b, rm, rv = (sd[pre + e].double() for e in (".weight", ".bias", ".running_mean", ".running_var"))
        return (t - rm.view(1, -1, 1, 1)) / torch.sqrt(rv.view(1, -1, 1, 1) + 1e-5) * w.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)

    w2 = {k: sd[k].double().requires_grad_() for k in names}
    x2 = x.double().requires_grad_()
    cur = x2
    for b in range(3):
        pre = "0.%d." % b
        m1, m2, m3 = (tok2map(m) for m in masks[b])
        o = bn(F.conv2d(cur, w2[pre + "conv1.weight"], stride=2 if b == 0 else 1), pre + "bn1") * m1
        o = bn(F.conv2d(o, w2[pre + "conv2.weight"], padding=1), pre + "bn2") * m2
        o = bn(F.conv2d(o, w2[pre + "conv3.weight"]), pre + "bn3")
        res = bn(F.conv2d(cur, w2[pre + "downsample.0.weight"], stride=2), pre + "downsample.1") if b == 0 else cur
        cur = (o + res) * m3
    ref2 = cur.mean(3).mean(2)
    ref2.backward(gy.double())
    assert rel(out.detach(), ref2.detach()) < 2e-3
    assert rel(xd.grad, x2.grad) < 2e-3, rel(xd.grad, x2.grad)
    for k in names:
        e = rel(got[k].grad, w2[k].grad)
        assert e < 2e-3, (k, e)


@pytest.mark.parametrize("G", [3, 10])
def test_sknet_training_matches_oracle_autograd(G):
    """SKBlock / SKNet forward + backward on the device (grouped tcgen05 GEMMs, im2col-free grouped wgrad) against
    fp64 autograd of the oracle restatement (blocks_coatt_transformer_sk.py:960-998); odd pair counts leave a ragged last
    128-row tile.  relu(z)**2 is C1, so -- unlike
    layer4 -- no ReLU-mask flips: tf32 arithmetic only, gate 3e-3 relative L2.  `fc` / `sk` get no gradient, as in the
    reference (its forward discards the selective-kernel attention)."""
    from ait_b200 import sk_train, synth
    from oracle import head_oracle
    head = synth.make_head(seed=0, calibrated=True, randomize_bn=True)
    sk = head.sk
    with torch.no_grad():   # non-zero biases (the reference initialises them to 0)
        for blk in (sk.sk_props, sk.sk_query):
            for c in (blk.convs[0][0], blk.convs[1][0]):
                c.bias.copy_(0.1 * torch.randn(1024, generator=torch.Generator().manual_seed(5)))
    sd = {k: v.clone() for k, v in sk.state_dict().items()}
    g = torch.Generator().manual_seed(40 + G)
    xp = torch.randn(G, 1024, 8, 8, generator=g).relu()
    xq = torch.randn(2, 1024, 8, 8, generator=g).relu()
    gp = torch.randn(G, 1024, 8, 8, generator=g)
    gq = torch.randn(2, 1024, 8, 8, generator=g)
    sd64 = {k: v.double().requires_grad_() for k, v in sd.items()}
    xp64, xq64 = xp.double().requires_grad_(), xq.double().requires_grad_()
    rp, rq = head_oracle.sknet_forward(sd64, xp64, xq64, dtype=torch.float64)
    (rp * gp.double()).sum().add((rq * gq.double()).sum()).backward()
    sk = sk.to(DEV).train()
    xpd, xqd = xp.to(DEV).requires_grad_(), xq.to(DEV).requires_grad_()
    op, oq = sk_train.sknet_train(sk, xpd, xqd)
    torch.autograd.backward([op, oq], [gp.to(DEV), gq.to(DEV)])
    torch.cuda.synchronize()
    assert _l2rel(op, rp.detach()) < 2e-3 and _l2rel(oq, rq.detach()) < 2e-3
    assert _l2rel(xpd.grad, xp64.grad) < 3e-3, _l2rel(xpd.grad, xp64.grad)
    assert _l2rel(xqd.grad, xq64.grad) < 3e-3
    errs = {}
    for name, p in sk.named_parameters():
        if ".convs." in name:
            errs[name] = _l2rel(p.grad, sd64[name].grad)
        else:
            assert p.grad is None and sd64[name].grad is None, name
    print("sknet train errs", {k: "%.1e" % v for k, v in errs.items()})
    assert len(errs) == 8 and max(errs.values()) < 3e-3, errs


class _OracleROIAlign(torch.autograd.Function):
    """C-oracle ROIAlign forward / backward (oracle/oracle_ops.c) as an autograd node of the fp64 reference graph."""

    @staticmethod
    def forward(ctx, feat, rois):
        from oracle import c_ops
        ctx.rois, ctx.shape = rois, feat.shape
        return torch.from_numpy(c_ops.roi_align_forward(feat.float().numpy(), rois.float().numpy(), 1.0 / 16.0, 7, 7, 0)).double()

    @staticmethod
    def backward(ctx, g):
        from oracle import c_ops
        B, Cc, H, W = ctx.shape
        d = c_ops.roi_align_backward(g.float().contiguous().numpy(), ctx.rois.float().numpy(), 1.0 / 16.0, 7, 7, B, Cc, H, W, 0)
        return torch.from_numpy(d).double(), None


def test_whole_head_training_step_matches_oracle_autograd():
    """DetectionHead.training_losses (ROIAlign -> AIT -> SKNet -> layer4 x2 -> heads -> the three RCNN losses), forward
    and backward on the device, against fp64 autograd over the oracle restatement of the same slice
    (faster_rcnn_coatt_transformer_sk.py:273-361).  Gates: losses 2e-3 relative; every gradient (both inputs, AIT 46,
    SK 8, layer4 10, heads 6) within 1e-1 relative L2 and cosine > 0.995 -- the tf32 forward flips the ReLU masks of
    layer4 / the FFNs for pre-activations within tf32 error of zero (see the layer-4 test: 4.7 % at its input by itself;
    the arithmetic with equal masks is 4-8e-4).  Second half: the SAME 72 gradients against the fp64 graph rebuilt with
    the device's own ReLU decisions (FFN masks from the AIT step's saved hidden tensors, layer-4 masks from both
    `_head_to_tail` calls) -- gate 4e-3 relative L2 (measured worst 1.3e-3): the whole gap of the first half is mask
    flips, the arithmetic of the chain is tf32-accurate."""
    import torch.nn.functional as F
    from ait_b200 import synth
    from oracle import head_oracle, target_oracle
    B, P = 2, 4
    head = synth.make_head(seed=0, calibrated=True, randomize_bn=True)
    from ait_b200.system.Models import set_dropout
    set_dropout(head, 0.0, 0.0)
    sd = {k: v.clone() for k, v in head.state_dict().items()}
    g = torch.Generator().manual_seed(71)
    maps = torch.stack([synth.c4_map(u) for u in range(B)])
    qrys = torch.stack([synth.query_feat(u) for u in range(B)])
    rois = torch.stack([synth.random_rois(u, P, batch_index=u) for u in range(B)])
    label = torch.tensor([[1, 0, 0, 1], [0, 0, 1, 0]]).view(-1)
    tgt = 0.3 * torch.randn(B * P, 4, generator=g)
    inw = (label > 0).float().view(-1, 1).expand(-1, 4).contiguous()
    outw = inw.clone()
    # fp64 reference graph
    pnames = {n for n, _ in head.named_parameters()}
    sd64 = {k: (v.double().requires_grad_() if k in pnames else (v.double() if v.is_floating_point() else v)) for k, v in sd.items()}
    m64, q64 = maps.double().requires_grad_(), qrys.double().requires_grad_()
    ref = head_oracle.head_forward(sd64, m64, q64, rois, dtype=torch.float64,
                                   roi_align_fn=lambda f, r: _OracleROIAlign.apply(f, r))
    ref_losses = target_oracle.rcnn_losses(ref["score"], ref["bbox_pred"].view(-1, 4), label, tgt.double(), inw.double(),
                                           outw.double(), B)
    sum(ref_losses).backward()
    # device (keeping what the second half of the test needs: the activations that carry the device's ReLU decisions)
    from ait_b200 import _lib as L, top_train
    from ait_b200.system import Models as ait_models
    top_train._saved_log_for_tests = []
    ait_models._keep_saved_for_tests[0] = True
    head = head.to(DEV).train()
    md, qd = maps.to(DEV).requires_grad_(), qrys.to(DEV).requires_grad_()
    try:
        losses = head.training_losses(md, qd, rois.to(DEV), label.to(DEV), tgt.to(DEV), inw.to(DEV), outw.to(DEV))
        l4_log, ait_saved = top_train._saved_log_for_tests, ait_models._keep_saved_for_tests[1]
    finally:
        top_train._saved_log_for_tests = None
        ait_models._keep_saved_for_tests[:] = [False, None]
    sum(losses).backward()
    torch.cuda.synchronize()
    for a, b in zip(losses, ref_losses):
        a, b = float(a.detach()), float(b.detach())
        assert abs(a - b) <= 2e-3 * max(abs(b), 1e-3), (a, b)

    def cos(a, b):
        a = a.detach().double().cpu().reshape(-1)
        return float(torch.dot(a, b.reshape(-1)) / (a.norm() * b.norm()))

    errs = {"non_img": (_l2rel(md.grad, m64.grad), cos(md.grad, m64.grad)),
            "non_qry": (_l2rel(qd.grad, q64.grad), cos(qd.grad, q64.grad))}
    n_none = 0
    for name, p in head.named_parameters():
        r = sd64[name].grad
        if ".bn" in name or "downsample.1" in name or re.match(r"sk\.sk_(props|query)\.(fc|sk)\.", name):
            assert p.grad is None, name      # frozen BatchNorm; the SK attention the reference discards
            n_none += 1
            continue
        assert p.grad is not None and r is not None, name
        errs[name] = (_l2rel(p.grad, r), cos(p.grad, r))
    worst = sorted(errs.items(), key=lambda kv: -kv[1][0])[:6]
    print("whole-head train: %d gradients, worst rel-L2 / cos:" % len(errs), [(k, "%.1e" % e, "%.4f" % c) for k, (e, c) in worst])
    assert len(errs) == 2 + 46 + 8 + 10 + 6, len(errs)
    bad = {k: v for k, v in errs.items() if v[0] > 1e-1 or v[1] < 0.995}
    assert not bad, bad

    # ---- the same comparison with the DEVICE's ReLU decisions injected into the fp64 graph (prove that
    # the gap above is mask flips, not arithmetic): the encoder / decoder FFN masks come out of the AIT step's saved hidden
    # tensors, the layer-4 masks out of both `_head_to_tail` calls' saved activations.  With equal masks both sides
    # differentiate the same piecewise-linear function and every gradient must agree to tf32 accuracy.
    lib = L.load()
    bp = B * P
    sv = ait_saved.view(torch.uint8)
    relu_masks = {}
    for tag, which in (("enc_ffn", 2), ("dec_ffn", 3)):
        off = int(lib.aitb_ait_saved_offset(B, P, which))
        hid = sv[off:off + bp * 64 * 2048 * 4].view(torch.float32).view(bp, 64, 2048)
        relu_masks[tag] = (hid > 0).double().cpu()
    assert len(l4_log) == 2                       # props call, then query call
    for call, saved in enumerate(l4_log):
        G = bp if call == 0 else B
        for b, blk in enumerate(saved):
            for j, t in enumerate(blk):
                relu_masks["l4.%d.%d.%d" % (call, b, j + 1)] = (t > 0).double().cpu().view(G, 4, 4, -1).permute(0, 3, 1, 2)
    used = set()

    def hook(tag, x):
        used.add(tag)
        m = relu_masks[tag]
        assert m.shape == x.shape, (tag, m.shape, x.shape)
        return m

    sd64b = {k: (v.double().requires_grad_() if k in pnames else (v.double() if v.is_floating_point() else v)) for k, v in sd.items()}
    m64b, q64b = maps.double().requires_grad_(), qrys.double().requires_grad_()
    head_oracle.RELU_HOOK = hook
    try:
        refb = head_oracle.head_forward(sd64b, m64b, q64b, rois, dtype=torch.float64,
                                        roi_align_fn=lambda f, r: _OracleROIAlign.apply(f, r))
    finally:
        head_oracle.RELU_HOOK = None
    assert used == set(relu_masks), used ^ set(relu_masks)
    ref_losses_b = target_oracle.rcnn_losses(refb["score"], refb["bbox_pred"].view(-1, 4), label, tgt.double(), inw.double(),
                                             outw.double(), B)
    sum(ref_losses_b).backward()
    errs_b = {"non_img": _l2rel(md.grad, m64b.grad), "non_qry": _l2rel(qd.grad, q64b.grad)}
    for name, p in head.named_parameters():
        if p.grad is not None:
            errs_b[name] = _l2rel(p.grad, sd64b[name].grad)
    worst_b = sorted(errs_b.items(), key=lambda kv: -kv[1])[:6]
    print("whole-head train, device ReLU masks injected: worst rel-L2:", [(k, "%.1e" % e) for k, e in worst_b])
    assert len(errs_b) == len(errs)
    bad_b = {k: v for k, v in errs_b.items() if v > 4e-3}          # measured worst 1.3e-3 (the plain comparison above: 3.7e-2)
    assert not bad_b, bad_b


@pytest.mark.parametrize("G,S,Cc,N,groups", [(5, 4, 512, 512, 1), (3, 8, 1024, 1024, 8), (1, 8, 256, 256, 2), (67, 4, 128, 128, 1)])
@pytest.mark.parametrize("taps", [1, 9])
def test_wgrad_conv_without_im2col(G, S, Cc, N, groups, taps):
    """aitb_wgrad_conv (X read through a shifted 4-D TMA view, zero-filled outside the map; n-tile = group) against
    (1) F.conv2d's weight gradient in fp64 on tf32-rounded operands and (2) the im2col + aitb_wgrad route it replaces
    (same products, same tensor-core path -> 1e-5).  Ragged: G*S*S not a multiple of the 32-row stages' 128-row splits."""
    import ctypes as C
    import torch.nn.functional as F
    from ait_b200 import _lib as L, ops
    g = torch.Generator().manual_seed(G * 100 + S + taps)
    x = _tf32(torch.randn(G, S, S, Cc, generator=g)).to(DEV)
    dy = _tf32(torch.randn(G * S * S, N, generator=g)).to(DEV)
    cg = Cc // groups
    dw = ops.wgrad_conv(dy, x.view(G * S * S, Cc), G, S, Cc, N, groups=groups, taps=taps)
    torch.cuda.synchronize()
    assert dw.shape == (N, taps * cg)
    k = 3 if taps == 9 else 1
    w = torch.zeros(N, cg, k, k, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(x.cpu().double().permute(0, 3, 1, 2), w, padding=k // 2, groups=groups)
    (y * dy.cpu().double().view(G, S, S, N).permute(0, 3, 1, 2)).sum().backward()
    ref = w.grad.permute(0, 2, 3, 1).reshape(N, taps * cg)           # tap-major
    assert _l2rel(dw, ref) < 2e-5, _l2rel(dw, ref)
    # accumulation into a given dW
    dw2 = ops.wgrad_conv(dy, x.view(G * S * S, Cc), G, S, Cc, N, groups=groups, taps=taps, dw=dw.clone())
    assert _l2rel(dw2, 2 * ref) < 2e-5
    if taps == 9:                                                     # the replaced route
        cols = torch.empty((G * S * S, groups, 9 * cg), dtype=torch.float32, device=DEV)
        L.check(L.load().aitb_im2col3x3_grouped(L.ptr(x), G, S, Cc, cg, L.ptr(cols), L.stream_ptr()))
        old = torch.zeros_like(dw)
        npg = N // groups
        for gi in range(groups):
            ops.wgrad(dy[:, gi * npg:(gi + 1) * npg], cols[:, gi], dw=old[gi * npg:(gi + 1) * npg], N=npg, K=9 * cg)
        assert _l2rel(dw, old.cpu().double()) < 1e-5


def test_head_training_loop_reduces_the_loss():
    """End to end on the device (tools/train_demo.py): device-side ProposalTargetLayer sample -> DetectionHead.training_losses
    -> backward -> plain torch.optim.SGD (lr 1e-2, no momentum) over the head's own nn.Parameters with the reference's
    stock init, eight steps on one fixed sample.  Gradient descent with a small step must go downhill: the total loss falls
    at EVERY step (measured 2.056 -> 2.007, classification 0.697 -> 0.660, box regression 0.571 -> 0.559) -- the gradients
    point downhill through the whole chain (ROIAlign, AIT, SKNet, layer4, heads, losses)."""
    import math
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import train_demo
    hist = train_demo.run(B=2, steps=8, lr=1e-2, dev=DEV, verbose=True, mom=0.0)
    tot = [sum(h) for h in hist]
    assert all(math.isfinite(t) for t in tot), tot
    assert all(b < a + 1e-4 for a, b in zip(tot, tot[1:])), tot
    assert tot[-1] < tot[0] - 0.03, tot
    assert hist[-1][0] < hist[0][0] - 0.02 and hist[-1][2] < hist[0][2] - 0.005, hist


def test_whole_head_training_step_matches_reference_golden_gradients():
    """tests/golden/head_grad.pt: losses and gradients of the UNMODIFIED reference modules (CPU fp32 autograd,
    tests/golden/make_golden_head_grad.py) on the inputs of the whole-head test above.  The device step (ROIAlign included,
    tf32 tensor-core math) must reproduce the three losses to 2e-3 relative and every one of the 70 parameter gradients and
    the query-feature gradient to 1e-1 relative L2 on the committed strided samples (ReLU-mask flips of a tf32 forward, see
    above), with norms within 10 %."""
    from conftest import load_golden
    from ait_b200 import synth
    gold = load_golden("head_grad.pt")
    B, P = gold["B"], gold["P"]
    head = synth.make_head(seed=0, calibrated=True, randomize_bn=True)
    from ait_b200.system.Models import set_dropout
    set_dropout(head, 0.0, 0.0)
    g = torch.Generator().manual_seed(gold["seed"])
    maps = torch.stack([synth.c4_map(u) for u in range(B)])
    qrys = torch.stack([synth.query_feat(u) for u in range(B)])
    rois = torch.stack([synth.random_rois(u, P, batch_index=u) for u in range(B)])
    label = torch.tensor([[1, 0, 0, 1], [0, 0, 1, 0]]).view(-1)
    tgt = 0.3 * torch.randn(B * P, 4, generator=g)
    inw = (label > 0).float().view(-1, 1).expand(-1, 4).contiguous()
    head = head.to(DEV).train()
    md, qd = maps.to(DEV).requires_grad_(), qrys.to(DEV).requires_grad_()
    losses = head.training_losses(md, qd, rois.to(DEV), label.to(DEV), tgt.to(DEV), inw.to(DEV), inw.to(DEV))
    sum(losses).backward()
    torch.cuda.synchronize()
    for a, b in zip(losses, gold["losses"]):
        a = float(a.detach())
        assert abs(a - b) <= 2e-3 * max(abs(b), 1e-3), (a, b)

    def check(name, grad, entry):
        assert grad is not None, name
        f = grad.detach().double().cpu().reshape(-1)
        s = f[::entry["stride"]][:entry["sample"].numel()]
        r = entry["sample"].double()
        e = float((s - r).norm() / r.norm())
        assert e < 1e-1, (name, e)
        assert abs(float(f.norm()) / entry["norm"] - 1.0) < 1e-1, name
        return e

    errs = {"non_qry": check("non_qry", qd.grad, gold["grad_query"])}
    params = dict(head.named_parameters())
    assert len(gold["params"]) == 70
    for name, entry in gold["params"].items():
        errs[name] = check(name, params[name].grad, entry)
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:4]
    print("whole-head train vs reference golden: worst rel-L2", [(k, "%.1e" % v) for k, v in worst])


def test_whole_head_training_step_with_default_dropout():
    """`DetectionHead().train()` out of the box (dropout 0.1 like `_fasterRCNN`'s Transformer, attention dropout 0.1): the
    whole-head training step runs with the masks active -- finite losses and gradients for every trainable parameter, a
    reproducible step under torch.manual_seed, and losses that differ from the dropout-free step (train mode no
    longer raises out of the box)."""
    from ait_b200 import synth
    from ait_b200.system.Models import set_dropout
    B, P = 2, 4
    head = synth.make_head(seed=0, calibrated=True, randomize_bn=True).to(DEV).train()
    maps = torch.stack([synth.c4_map(u) for u in range(B)]).to(DEV).requires_grad_()
    qrys = torch.stack([synth.query_feat(u) for u in range(B)]).to(DEV).requires_grad_()
    rois = torch.stack([synth.random_rois(u, P, batch_index=u) for u in range(B)]).to(DEV)
    g = torch.Generator().manual_seed(3)
    label = (torch.rand(B * P, generator=g) < 0.5).long().to(DEV)
    tgt = (0.3 * torch.randn(B * P, 4, generator=g)).to(DEV)
    inw = (label > 0).float().view(-1, 1).expand(-1, 4).contiguous()

    def step(seed):
        head.zero_grad(set_to_none=True)
        maps.grad = None
        torch.manual_seed(seed)
        losses = head.training_losses(maps, qrys, rois, label, tgt, inw, inw)
        sum(losses).backward()
        torch.cuda.synchronize()
        return [float(x.detach()) for x in losses], maps.grad.clone()

    l1, g1 = step(7)
    assert head.transformer.last_dropout_seed != 0
    l2, g2 = step(7)
    l3, _ = step(8)
    assert all(torch.isfinite(torch.tensor(l1))) and torch.isfinite(g1).all()
    # no gradient by design: frozen BatchNorm (`set_bn_fix`) and the `sk.*.fc` / `sk.*.sk` layers SKBlock.forward discards
    frozen = lambda n: ".bn" in n or "downsample.1" in n or (n.startswith("sk.") and (".fc." in n or ".sk." in n))   # noqa: E731
    missing = [n for n, p_ in head.named_parameters() if p_.requires_grad and p_.grad is None and not frozen(n)]
    assert not missing, missing
    assert sum(1 for p_ in head.parameters() if p_.grad is not None) == 70
    assert l1 == l2                                     # same seed, same masks (losses are reduced deterministically)
    assert float((g1 - g2).abs().max()) <= 1e-3 * float(g1.abs().max())       # ROIAlign backward: atomics order only
    assert l1 != l3
    set_dropout(head, 0.0, 0.0)
    l0, _ = step(7)
    assert head.transformer.last_dropout_seed == 0 and l0 != l1


def test_detector_tail_training_step_end_to_end():
    """The reference's whole training forward after the backbone (faster_rcnn_coatt_transformer_sk.py:229-361) on the device:
    co-attention -> RPN (+ anchor targets, RPN losses) -> proposal targets -> ROIAlign -> AIT (dropout active) -> SKNet -> layer4
    -> heads -> RCNN losses; one backward fills the gradient of the backbone features and of every trainable parameter of every
    stage; three SGD steps on one batch run through (losses finite and stable)."""
    from ait_b200 import synth
    from ait_b200.detector import DetectorTail
    from ait_b200.targets import ProposalTargetLayer
    B, H, W = 2, 19, 31
    g = torch.Generator().manual_seed(12)
    torch.manual_seed(1)
    m = DetectorTail(rpn_cfg={"TEST": dict(pre_nms_topN=6000, post_nms_topN=300, nms_thresh=0.7),
                              "TRAIN": dict(pre_nms_topN=2000, post_nms_topN=64, nms_thresh=0.7)})
    with torch.no_grad():                       # the stock init zeroes the GroupNorm affine: give the non-local branch a gradient path
        for seq in (m.coattention_module.coattention.theta, m.coattention_module.coattention.omega):
            seq[1].weight.fill_(0.5)
    m = m.to(DEV).train()
    for n, p in m.named_parameters():
        if ".bn" in n or "downsample.1" in n:
            p.requires_grad_(False)             # frozen BatchNorm (set_bn_fix)
    img = (torch.randn(B, 1024, H, W, generator=g).relu() * 0.5).to(DEV).requires_grad_()
    qry = (torch.randn(B, 1024, 8, 8, generator=g).relu() * 0.5).to(DEV).requires_grad_()
    im_info = torch.tensor([[300.0, 500.0, 1.0], [300.0, 500.0, 1.0]], device=DEV)
    gt = torch.zeros(B, 4, 5, device=DEV)
    gt[0, 0] = torch.tensor([40.0, 30.0, 200.0, 180.0, 1.0])
    gt[0, 1] = torch.tensor([250.0, 100.0, 420.0, 260.0, 1.0])
    gt[1, 0] = torch.tensor([60.0, 50.0, 300.0, 220.0, 1.0])
    nb = torch.tensor([2, 1], device=DEV)
    sampler = ProposalTargetLayer(2, cfg={"BATCH_SIZE": 32}, rng="device", seed=3)
    opt = torch.optim.SGD([p for p in m.parameters() if p.requires_grad], lr=2e-3)
    totals = []
    for it in range(3):
        opt.zero_grad(set_to_none=True)
        img.grad = None
        qry.grad = None
        np.random.seed(4)                        # the anchor sampler: the same anchors every step
        torch.manual_seed(7)                     # the dropout seed: the same masks every step
        sampler._calls = 0                       # the proposal sampler: the same draw every step
        rois, l_rc, l_rb, l_c, l_m, l_b, label = m.training_step(img, qry, im_info, gt, nb, sampler=sampler)
        total = l_rc + l_rb + l_c + l_m + l_b
        total.backward()
        totals.append(float(total.detach()))
        if it == 0:
            assert rois.shape == (B, 32, 5) and label.shape == (B * 32,)
            assert all(bool(torch.isfinite(x.detach()).all()) for x in (l_rc, l_rb, l_c, l_m, l_b))
            assert bool(torch.isfinite(img.grad).all()) and float(img.grad.abs().max()) > 0
            assert bool(torch.isfinite(qry.grad).all()) and float(qry.grad.abs().max()) > 0
            dead = ("sk.sk_props.fc", "sk.sk_props.sk", "sk.sk_query.fc", "sk.sk_query.sk")
            missing = [n for n, p in m.named_parameters()
                       if p.requires_grad and not n.startswith(dead) and (p.grad is None or not bool(torch.isfinite(p.grad).all()))]
            assert not missing, missing
            for prefix in ("coattention_module.", "RCNN_rpn.", "transformer.", "sk.", "RCNN_top.", "RCNN_cls_score.", "RCNN_bbox_pred."):
                gsum = sum(float(p.grad.abs().sum()) for n, p in m.named_parameters() if n.startswith(prefix) and p.grad is not None)
                assert gsum > 0, prefix
        opt.step()
    print("detector tail training totals", totals)
    # the proposals (hence the sampled rois and the RCNN losses) move with the RPN weights from step to step, so the total is not
    # a fixed objective: it must stay finite and in the neighbourhood of the first step (measured 2.213, 2.227, 2.182); the
    # fixed-sample descent check is test_head_training_loop_reduces_the_loss
    assert all(np.isfinite(totals)) and max(abs(t - totals[0]) for t in totals) < 0.5, totals
