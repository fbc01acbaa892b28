"""GPU parity of the whole head (ROIAlign -> AIT -> SKNet -> RCNN_top -> heads) against the golden
vectors produced by the reference and against the CPU oracle; module-level drop-ins; size-independent
properties at the benchmark configuration."""
import pytest
import torch

from conftest import golden_head, head_inputs, load_golden
from oracle import head_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

# Stated tolerances.  north_star: cls_prob within 1e-3 absolute in the fp32 configuration, bf16 reported
# separately.  Three compute configurations exist (include/aitb200.h):
#   "fp32"  split bf16 hi/lo planes, three tensor-core passes per product: THE fp32 configuration; it meets
#           the 1e-3 gate even on the stress test below (score layer fitted to spread cls_prob over (0.07, 0.99))
#   "tf32"  fp32 storage + tf32 math: 10-bit operand mantissa, measured feature error 4.7e-4 of scale
#           -> 5e-3 on the stress-test cls_prob, gate 1e-2 (reported separately, like bf16)
#   "bf16"  3.5e-3 feature error -> 1e-2 on cls_prob, gate 3e-2
# With the reference's own random init the scores are degenerate (0.0057 +- 1e-6) and
# every configuration meets 1e-3 with orders of magnitude to spare (test_head_stock_init_scores).
CLS_ATOL = {"fp32": 1e-3, "tf32": 1e-2, "bf16": 3e-2}
# intermediates: max |err| relative to the tensor's own scale (max |ref|)
REL = {"fp32": 2e-4, "tf32": 4e-3, "bf16": 4e-2}
MODES = ["fp32", "tf32", "bf16"]


def _scaled_err(out, ref):
    return float((out.double() - ref.double()).abs().max() / ref.double().abs().max().clamp_min(1e-12))


@pytest.mark.parametrize("dtype", MODES + ["fp32-three-pass"])
def test_head_matches_reference_golden(dtype, monkeypatch):
    if dtype == "fp32-three-pass":        # the fp32 configuration without the precision plan (A/B switch)
        monkeypatch.setenv("AITB_PRECISION_PLAN", "0")
        dtype = "fp32"
    head, g = golden_head(compute_dtype=dtype)
    head = head.to(DEV)
    non_img, non_qry, rois = head_inputs(g["B"], g["P"])
    cls_prob, bbox, taps = head(non_img.to(DEV), non_qry.to(DEV), rois.to(DEV), taps=True)
    bp = g["B"] * g["P"]
    pooled = taps["pooled"].float().cpu().permute(0, 2, 1).reshape(bp, 1024, 7, 7)
    ait = taps["ait_out"].float().cpu().permute(0, 2, 1).reshape(bp, 1024, 8, 8)
    sk = taps["sk_out"].float().cpu().permute(0, 2, 1).reshape(bp, 1024, 8, 8)
    if dtype in ("tf32", "fp32"):
        # the north-star ROIAlign gate (1e-5 relative) INSIDE the head.  tf32: fp32 storage, the tap is the exact ROIAlign
        # output.  fp32: the pooled features are stored as two bf16 planes hi = bf16(x), lo = bf16(x - hi); |x - hi - lo|
        # <= 2^-9 |x - hi| <= 2^-17 |x| = 7.6e-6 |x| -- the storage format itself stays inside the gate, asserted here.
        torch.testing.assert_close(pooled[:, ::16], g["pooled_s"], rtol=1e-5, atol=1e-6)   # ROIAlign gate
    assert _scaled_err(pooled[:, ::16], g["pooled_s"]) < REL[dtype]
    assert _scaled_err(ait[:, ::16], g["ait_s"]) < REL[dtype]
    assert _scaled_err(sk[:, ::16], g["sk_s"]) < 2 * REL[dtype]
    assert _scaled_err(taps["feat"].cpu(), g["feat"]) < 2 * REL[dtype]
    assert _scaled_err(taps["qfeat"].cpu(), g["qfeat"]) < 2 * REL[dtype]
    assert _scaled_err(bbox.cpu(), g["bbox_pred"]) < 4 * REL[dtype]
    torch.testing.assert_close(cls_prob.cpu(), g["cls_prob"], rtol=0, atol=CLS_ATOL[dtype])


# enc_out sits INSIDE the one-pass region of the fp32 configuration's precision plan (encoder GEMMs on 11-bit fp16 hi planes,
# DESIGN.md): tf32-class at that tap by design (measured 3.4e-4 of scale); every tensor downstream of the decoder -- ait_out,
# sk_out, the layer-4 features, bbox_pred, cls_prob -- keeps its gate.  With the plan off (three passes everywhere) enc_out
# meets REL["fp32"] too.
ENC_REL = {"plan": 6e-4, "three-pass": REL["fp32"]}


@pytest.mark.parametrize("plan", ["plan", "three-pass"])
@pytest.mark.parametrize("B,P", [(1, 1), (3, 5), (2, 37)])
def test_head_matches_oracle_ragged_sizes(B, P, plan, monkeypatch):
    """sizes that do not fill the 128-row GEMM tiles / straddle pairs and units."""
    if plan == "three-pass":
        monkeypatch.setenv("AITB_PRECISION_PLAN", "0")
    head, _ = golden_head()
    sd = head.state_dict()
    head = head.to(DEV)
    assert head.engine().plan == (1 if plan == "plan" else 0)
    non_img, non_qry, rois = head_inputs(B, P, first_unit=10)
    with torch.no_grad():
        ref = head_oracle.head_forward(sd, non_img, non_qry, rois)
    cls_prob, bbox, taps = head(non_img.to(DEV), non_qry.to(DEV), rois.to(DEV), taps=True)
    enc = taps["enc_out"].float().cpu()
    assert _scaled_err(enc[:, :49], ref["enc_out"][:, :49]) < ENC_REL[plan]    # pad rows are dead after enc self-attn
    ait = taps["ait_out"].float().cpu().permute(0, 2, 1).reshape(B * P, 1024, 8, 8)
    assert _scaled_err(ait, ref["ait_out"]) < REL["fp32"]
    assert _scaled_err(taps["feat"].cpu(), ref["feat"]) < 2 * REL["fp32"]
    torch.testing.assert_close(cls_prob.cpu(), ref["cls_prob"], rtol=0, atol=CLS_ATOL["fp32"])
    assert _scaled_err(bbox.cpu(), ref["bbox_pred"]) < 4 * REL["fp32"]


@pytest.mark.parametrize("dtype,atol", [("fp32", 1e-3), ("tf32", 1e-3), ("bf16", 1e-3)])
def test_head_stock_init_scores(dtype, atol):
    """The north-star gate as written: random-init head weights (the reference's `_init_weights`),
    cls_prob within 1e-3 absolute of the reference implementation."""
    from ait_b200 import synth
    head = synth.make_head(seed=3, compute_dtype=dtype)
    sd = {k: v.clone() for k, v in head.state_dict().items()}
    head = head.to(DEV)
    non_img, non_qry, rois = head_inputs(2, 6, first_unit=20)
    with torch.no_grad():
        ref = head_oracle.head_forward(sd, non_img, non_qry, rois)
    cls_prob, bbox = head(non_img.to(DEV), non_qry.to(DEV), rois.to(DEV))
    torch.testing.assert_close(cls_prob.cpu(), ref["cls_prob"], rtol=0, atol=atol)
    assert _scaled_err(bbox.cpu(), ref["bbox_pred"]) < 4 * REL[dtype]


def test_transformer_module_drop_in_matches_reference_golden():
    """adaptive_image_transformer.py usage: Transformer(...)(x_props=..., x_query=...)."""
    from ait_b200.system.Models import Transformer
    head, _ = golden_head()
    g = load_golden("ait_rand.pt")
    t = Transformer(d_k=64, d_v=64, d_model=512, d_word_vec=512, d_inner=2048, n_position=64, n_layers=1, n_head=8,
                    dropout=0.1)
    t.load_state_dict(head.transformer.state_dict(), strict=True)
    t = t.to(DEV).eval()
    gen = torch.Generator().manual_seed(g["seed"])
    xp = torch.rand(6, 1024, 7, 7, generator=gen)
    xq = torch.rand(2, 1024, 8, 8, generator=gen)
    out = t(x_props=xp.to(DEV), x_query=xq.to(DEV))
    assert out.shape == (6, 1024, 8, 8) and out.dtype == torch.float32
    assert _scaled_err(out.cpu()[:, ::8], g["out_s"]) < REL["fp32"]
    # .train(): dropout is neither skipped nor refused -- the training step draws a fresh seed per call from torch's generator
    t.train()
    torch.manual_seed(5)
    o1 = t(x_props=xp.to(DEV), x_query=xq.to(DEV))
    s1 = t.last_dropout_seed
    o2 = t(x_props=xp.to(DEV), x_query=xq.to(DEV))
    torch.manual_seed(5)
    o3 = t(x_props=xp.to(DEV), x_query=xq.to(DEV))
    assert s1 != 0 and t.last_dropout_seed == s1
    assert torch.equal(o1, o3) and not torch.equal(o1, o2)
    assert float((o1.detach() - out).abs().max()) > 1e-2 * float(out.abs().max())     # p = 0.1 masks move the output visibly


def test_sknet_and_head_to_tail_modules_match_oracle():
    head, _ = golden_head()
    sd = head.state_dict()
    head = head.to(DEV)
    g = torch.Generator().manual_seed(21)
    xp = torch.randn(5, 1024, 8, 8, generator=g)
    xq = torch.relu(torch.randn(2, 1024, 8, 8, generator=g))
    sp, sq = head.sk(xp.to(DEV), xq.to(DEV))
    with torch.no_grad():
        rp, rq = head_oracle.sknet_forward({k[3:]: v for k, v in sd.items() if k.startswith("sk.")}, xp, xq)
        rf = head_oracle.head_to_tail({k[9:]: v for k, v in sd.items() if k.startswith("RCNN_top.")}, rp)
    assert _scaled_err(sp.cpu(), rp) < REL["fp32"] and _scaled_err(sq.cpu(), rq) < REL["fp32"]
    feat = head.engine().top_forward(rp.to(DEV))
    assert _scaled_err(feat.cpu(), rf) < REL["fp32"]


def test_benchmark_shape_properties():
    """config 2 (8 units x 300 proposals): results do not depend on how units are batched (every output
    row is one MMA accumulation chain), proposals are permutation-equivariant, outputs are finite."""
    from ait_b200 import synth
    from ait_b200.proposal import propose_rois
    head = synth.make_head(seed=0, calibrated=True, randomize_bn=True).to(DEV)
    B, P = 8, 300
    non_img = torch.stack([synth.c4_map(u) for u in range(B)]).to(DEV)
    non_qry = torch.stack([synth.query_feat(u) for u in range(B)]).to(DEV)
    data = [synth.rpn_outputs(u) for u in range(B)]
    rois, n_keep = propose_rois(torch.stack([d[0] for d in data]).to(DEV), torch.stack([d[1] for d in data]).to(DEV))
    assert n_keep.tolist() == [P] * B
    cls_all, bbox_all = head(non_img, non_qry, rois)
    assert torch.isfinite(cls_all).all() and torch.isfinite(bbox_all).all()
    assert float(cls_all.min()) >= 0 and float(cls_all.max()) <= 1
    # unit 5 alone == unit 5 inside the batch, bit for bit
    r5 = rois[5:6].clone()
    r5[..., 0] = 0
    c5, b5 = head(non_img[5:6], non_qry[5:6], r5)
    assert torch.equal(c5[0], cls_all[5]) and torch.equal(b5[0], bbox_all[5])
    # permuting the proposals of a unit permutes its outputs
    perm = torch.randperm(P, generator=torch.Generator().manual_seed(0)).to(DEV)
    cp, bpred = head(non_img[5:6], non_qry[5:6], r5[:, perm])
    assert torch.equal(cp[0], c5[0][perm]) and torch.equal(bpred[0], b5[0][perm])


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_benchmark_shape_units_match_oracle(dtype):
    """BASELINE configs[1] at full size (8 units x 300 proposals, NMS-produced rois): two whole units of the batch are
    compared with the CPU oracle -- rois bit-exact, cls_prob on the spread-calibrated score layer within the stated
    tolerance (1e-3 for the fp32 configuration), layer-4 features and bbox_pred relative to their scale."""
    from ait_b200 import synth
    from ait_b200.proposal import propose_rois
    head = synth.spread_score_layer(synth.make_head(seed=0, calibrated=True, randomize_bn=True, compute_dtype=dtype))
    sd = {k: v.clone() for k, v in head.state_dict().items()}
    head = head.to(DEV)
    B, P = 8, 300
    non_img = torch.stack([synth.c4_map(u) for u in range(B)])
    non_qry = torch.stack([synth.query_feat(u) for u in range(B)])
    data = [synth.rpn_outputs(u) for u in range(B)]
    boxes, scores = torch.stack([d[0] for d in data]), torch.stack([d[1] for d in data])
    rois, n_keep = propose_rois(boxes.to(DEV), scores.to(DEV))
    cls_all, bbox_all, taps = head(non_img.to(DEV), non_qry.to(DEV), rois, taps=True)
    feat_all = taps["feat"].cpu().view(B, P, 2048)
    spans, lo, hi = [], 1.0, 0.0
    for u in (0, 1):
        ref_rois, _ = head_oracle.propose_rois(boxes[u:u + 1], scores[u:u + 1], 6000, P, 0.7)
        mine = rois[u:u + 1].cpu().clone()
        mine[..., 0] = 0
        assert torch.equal(mine, ref_rois)
        with torch.no_grad():
            ref = head_oracle.head_forward(sd, non_img[u:u + 1], non_qry[u:u + 1], ref_rois)
        spans.append(float(ref["cls_prob"].max() - ref["cls_prob"].min()))
        lo, hi = min(lo, float(ref["cls_prob"].min())), max(hi, float(ref["cls_prob"].max()))
        torch.testing.assert_close(cls_all[u:u + 1].cpu(), ref["cls_prob"], rtol=0, atol=CLS_ATOL[dtype])
        assert _scaled_err(feat_all[u], ref["feat"]) < 2 * REL[dtype]
        assert _scaled_err(bbox_all[u:u + 1].cpu(), ref["bbox_pred"]) < 4 * REL[dtype]
    # the score gate is not vacuous: 0.03 .. 0.30 inside unit 0, 0.70 .. 0.99 inside unit 1 (oracle, measured)
    assert min(spans) > 0.2 and hi - lo > 0.5, (spans, lo, hi)


def test_unit_chunking_is_invisible():
    """More units than one library call takes (workspace bound): identical to running the chunks by hand."""
    head, _ = golden_head()
    head = head.to(DEV)
    eng = head.engine()
    B, P = 5, 6
    non_img, non_qry, rois = head_inputs(B, P, first_unit=30)
    non_img, non_qry, rois = non_img.to(DEV), non_qry.to(DEV), rois.to(DEV)
    ref_c, ref_b = eng.head_forward(non_img, non_qry, rois)
    old = eng.MAX_UNITS_PER_CALL
    try:
        eng.MAX_UNITS_PER_CALL = 2
        c, b = eng.head_forward(non_img, non_qry, rois)
    finally:
        eng.MAX_UNITS_PER_CALL = old
    assert torch.equal(c, ref_c) and torch.equal(b, ref_b)


def test_packed_weight_cache_follows_in_place_updates():
    """eval forward -> in-place parameter update (what optimizer.step() does) -> eval forward must serve the NEW
    weights: the packed engine is re-built when a parameter's version counter moves."""
    head, _ = golden_head()
    head = head.to(DEV)
    non_img, non_qry, rois = (t.to(DEV) for t in head_inputs(1, 3))
    c0, b0 = head(non_img, non_qry, rois)
    c0, b0 = c0.clone(), b0.clone()
    c_same, b_same = head(non_img, non_qry, rois)
    assert torch.equal(c_same, c0) and torch.equal(b_same, b0)
    eng0 = head._engine
    with torch.no_grad():
        head.RCNN_bbox_pred.weight.mul_(2.0)                       # in-place, like an optimizer step
        head.transformer.dec_trans[0].bias.add_(0.25)
    c1, b1 = head(non_img, non_qry, rois)
    assert head._engine is not eng0
    assert not torch.equal(b1, b0) and not torch.equal(c1, c0)
    sd = {k: v.detach().cpu().clone() for k, v in head.state_dict().items()}
    with torch.no_grad():
        ref = head_oracle.head_forward(sd, non_img.cpu(), non_qry.cpu(), rois.cpu())
    assert _scaled_err(b1.cpu(), ref["bbox_pred"]) < 4 * REL["fp32"]
    # the Transformer module's own cache
    t = head.transformer
    xp, xq = torch.rand(2, 1024, 7, 7, device=DEV), torch.rand(1, 1024, 8, 8, device=DEV)
    o0 = t(x_props=xp, x_query=xq).clone()
    with torch.no_grad():
        t.dec_trans[0].bias.add_(1.0)
    o1 = t(x_props=xp, x_query=xq)
    torch.testing.assert_close(o1, o0 + 1.0, rtol=2e-5, atol=2e-5)   # two-plane output: 2^-17 relative


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_modules_on_a_non_current_device():
    """a model living on cuda:1 while the current device is cuda:0 -- every wrapper launches on the device (and the
    current stream of the device) that owns its tensors."""
    from ait_b200.proposal import propose_rois
    from ait_b200.roi_layers import nms
    from ait_b200 import synth
    assert torch.cuda.current_device() == 0
    d1 = "cuda:1"
    head, _ = golden_head()
    sd = {k: v.clone() for k, v in head.state_dict().items()}
    head = head.to(d1)
    B, P = 2, 5
    non_img = torch.stack([synth.c4_map(u) for u in range(B)])
    non_qry = torch.stack([synth.query_feat(u) for u in range(B)])
    data = [synth.rpn_outputs(u) for u in range(B)]
    boxes, scores = torch.stack([d[0] for d in data]), torch.stack([d[1] for d in data])
    rois, _ = propose_rois(boxes.to(d1), scores.to(d1), 6000, P, 0.7)
    assert rois.device == torch.device(d1)
    ref_rois, _ = head_oracle.propose_rois(boxes, scores, 6000, P, 0.7)
    assert torch.equal(rois.cpu(), ref_rois)
    cls_prob, bbox = head(non_img.to(d1), non_qry.to(d1), rois)
    with torch.no_grad():
        ref = head_oracle.head_forward(sd, non_img, non_qry, ref_rois)
    torch.testing.assert_close(cls_prob.cpu(), ref["cls_prob"], rtol=0, atol=CLS_ATOL["fp32"])
    keep = nms(boxes[0, :500].to(d1), scores[0, :500].to(d1), 0.5)
    assert keep.device == torch.device(d1) and keep.numel() > 0
    assert torch.cuda.current_device() == 0


def test_pipeline_cuda_graph_replay_matches_eager():
    """DetectionPipeline(graph=True): proposal tail + head captured once per shape and replayed; identical results to the
    eager launches, also for new inputs of the same shape."""
    from ait_b200 import synth
    from ait_b200.pipeline import DetectionPipeline
    head, _ = golden_head()
    head = head.to(DEV)
    B, P = 2, 20
    pipe_g = DetectionPipeline(head, 6000, P, 0.7, graph=True)
    pipe_e = DetectionPipeline(head, 6000, P, 0.7, graph=False)
    for first in (0, 7, 0):
        non_img = torch.stack([synth.c4_map(first + u) for u in range(B)]).to(DEV)
        non_qry = torch.stack([synth.query_feat(first + u) for u in range(B)]).to(DEV)
        data = [synth.rpn_outputs(first + u) for u in range(B)]
        boxes, scores = torch.stack([d[0] for d in data]).to(DEV), torch.stack([d[1] for d in data]).to(DEV)
        re, ce, be = pipe_e(non_img, non_qry, boxes, scores)
        rg, cg, bg = pipe_g(non_img, non_qry, boxes, scores)
        torch.cuda.synchronize()
        assert torch.equal(rg, re) and torch.equal(cg, ce) and torch.equal(bg, be)
    assert len(pipe_g._captured) == 1
