"""GPU parity (through the C ABI): NMS / top-n / proposal tail (bit-exact) and ROIAlign fwd/bwd."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import c_ops, head_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rand_boxes(n, seed, span=400.0, size=150.0):
    g = torch.Generator().manual_seed(seed)
    xy = torch.rand(n, 2, generator=g) * span
    wh = torch.rand(n, 2, generator=g) * size + 1
    return torch.cat([xy, xy + wh], 1), torch.rand(n, generator=g)


@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 127, 129, 500, 3000])
@pytest.mark.parametrize("thr", [0.3, 0.7])
def test_nms_matches_oracle_bit_exact(n, thr):
    from ait_b200.roi_layers import nms
    boxes, scores = _rand_boxes(n, 100 + n)
    keep = nms(boxes.to(DEV), scores.to(DEV), thr)
    assert keep.dtype == torch.int64 and keep.is_cuda
    ref = c_ops.nms(boxes.numpy(), scores.numpy(), thr, ge=False)
    assert np.array_equal(keep.cpu().numpy(), ref)


def test_nms_ties_and_edge_cases():
    from ait_b200.roi_layers import nms
    boxes = torch.tensor([[0, 0, 9, 9], [0, 0, 9, 4], [20, 20, 29, 29]], dtype=torch.float32)
    s = torch.tensor([0.9, 0.8, 0.7])
    assert nms(boxes.to(DEV), s.to(DEV), 0.5).tolist() == [0, 1, 2]          # IoU == thr is kept (nms.cu:60)
    assert nms(boxes.to(DEV), s.to(DEV), 0.49).tolist() == [0, 2]
    same = torch.tensor([[5, 5, 50, 50]], dtype=torch.float32).repeat(70, 1)
    assert nms(same.to(DEV), torch.linspace(1, 0, 70).to(DEV), 0.7).tolist() == [0]
    eq = torch.tensor([[0, 0, 10, 10], [0, 0, 10, 10], [100, 100, 110, 110]], dtype=torch.float32)
    assert nms(eq.to(DEV), torch.full((3,), 0.5).to(DEV), 0.7).tolist() == [0, 2]     # stable on equal scores
    e = nms(torch.zeros(0, 4, device=DEV), torch.zeros(0, device=DEV), 0.5)
    assert e.numel() == 0 and e.device.type == "cpu" and e.dtype == torch.int64
    with pytest.raises(RuntimeError):
        nms(torch.zeros(3, 4), torch.zeros(3), 0.5)                            # CPU tensors: no fallback


def test_nms_rpn_site_matches_reference_golden():
    """unit 0, VOC anchors: top-6000, thr 0.7 -- the keep list produced by the reference's own kernel."""
    from ait_b200 import ops, synth
    from ait_b200.roi_layers import nms
    g = load_golden("nms_rpn_unit0.pt")
    boxes, scores = synth.rpn_outputs(0)
    bd, sd = boxes.to(DEV), scores.to(DEV)
    order = ops.topk_desc(sd[None], 6000)
    ref_order = np.argsort(-scores.double().numpy(), kind="stable")[:6000]
    assert np.array_equal(order[0].cpu().numpy(), ref_order)
    keep = nms(bd[order[0]], sd[order[0]], 0.7)
    assert keep.numel() == g["n_keep"]
    assert torch.equal(keep[:300].cpu(), g["keep_first300"])
    keep_all = nms(bd, sd, 0.7)
    assert keep_all.numel() == g["n_keep_all"] and torch.equal(keep_all[:64].cpu(), g["keep_all_first64"])


def test_topk_with_ties_is_stable():
    from ait_b200 import ops
    g = torch.Generator().manual_seed(3)
    s = torch.randint(0, 50, (3, 5000), generator=g).float() / 7.0        # many exact ties
    s[1, :100] = -s[1, :100]
    for n in (1, 100, 4999, 5000):
        order = ops.topk_desc(s.to(DEV), n).cpu().numpy()
        for b in range(3):
            ref = np.argsort(-s[b].double().numpy(), kind="stable")[:n]
            assert np.array_equal(order[b], ref)


def test_topk_crowded_threshold_bin_spills_and_stays_exact():
    """all scores inside ONE bin of the 12-bit histogram (so every score is a candidate: 30 000 > the 16 384 the fused kernel keeps
    in shared memory -> bucket segments spill to the workspace), and a heavily tied distribution (a few distinct values: one bucket
    holds thousands of keys, ranks by counting)."""
    from ait_b200 import ops
    g = torch.Generator().manual_seed(5)
    s = 0.5 + 0.04 * torch.rand(2, 30000, generator=g)
    s[1] = (torch.randint(0, 4, (30000,), generator=g).float() + 1.0) / 8.0          # four distinct values
    for n in (6000, 20000, 30000):
        order = ops.topk_desc(s.to(DEV), n).cpu().numpy()
        for b in range(2):
            ref = np.argsort(-s[b].double().numpy(), kind="stable")[:n]
            assert np.array_equal(order[b], ref)


@pytest.mark.parametrize("scales,pre,post", [((8, 16, 32), 6000, 300), ((4, 8, 16, 32), 12000, 2000)])
def test_proposal_tail_matches_oracle_bit_exact(scales, pre, post):
    """rows a1+a2: rois [B, post, 5] identical to the reference loop (proposal_layer.py:129-166)."""
    from ait_b200 import synth
    from ait_b200.proposal import propose_rois
    B = 3
    data = [synth.rpn_outputs(u, scales=scales) for u in range(B)]
    boxes = torch.stack([d[0] for d in data])
    scores = torch.stack([d[1] for d in data])
    rois, n_keep = propose_rois(boxes.to(DEV), scores.to(DEV), pre, post, 0.7)
    ref, counts = head_oracle.propose_rois(boxes, scores, pre, post, 0.7)
    assert n_keep.tolist() == counts
    assert torch.equal(rois.cpu(), ref)
    if post == 300:
        assert torch.equal(rois[0].cpu(), load_golden("nms_rpn_unit0.pt")["rois"])


@pytest.mark.parametrize("pre,post", [(6000, 1), (6000, 5), (6000, 64), (1000, 65), (130, 300), (6000, 1024), (6000, 1025)])
def test_proposal_tail_early_stop_paths(pre, post):
    """post <= 1024 takes the kept-list kernel (tests only against boxes kept so far, stops at `post`);
    1025 takes the bitmask + scan path.  Both must reproduce the reference loop bit for bit."""
    from ait_b200 import synth
    from ait_b200.proposal import propose_rois
    data = [synth.rpn_outputs(40 + u) for u in range(2)]
    boxes = torch.stack([d[0] for d in data])
    scores = torch.stack([d[1] for d in data])
    rois, n_keep = propose_rois(boxes.to(DEV), scores.to(DEV), pre, post, 0.7)
    ref, counts = head_oracle.propose_rois(boxes, scores, pre, post, 0.7)
    assert n_keep.tolist() == counts
    assert torch.equal(rois.cpu(), ref)


def test_proposal_tail_zero_padding_when_few_survive():
    from ait_b200.proposal import propose_rois
    boxes = torch.tensor([[10, 10, 50, 50]], dtype=torch.float32).repeat(200, 1)[None].repeat(2, 1, 1)
    boxes[1, 100:] += 300
    scores = torch.rand(2, 200, generator=torch.Generator().manual_seed(1))
    rois, n_keep = propose_rois(boxes.to(DEV), scores.to(DEV), 6000, 300, 0.7)
    assert n_keep.tolist() == [1, 2]
    r = rois.cpu()
    assert torch.all(r[0, 1:, 1:] == 0) and torch.all(r[1, 2:, 1:] == 0)
    assert torch.all(r[0, :, 0] == 0) and torch.all(r[1, :, 0] == 1)


# ---------------------------------------------------------------------------------------- ROIAlign
def test_roi_align_matches_reference_golden_edge_cases():
    from ait_b200.roi_layers import ROIAlign
    g = load_golden("roi_align_small.pt")
    feat = torch.randn(2, 8, 38, 63, generator=torch.Generator().manual_seed(g["seed"]))
    out = ROIAlign((7, 7), 1 / 16.0, 0)(feat.to(DEV), g["rois"].to(DEV)).cpu()
    torch.testing.assert_close(out, g["out"], rtol=1e-5, atol=1e-6)
    out2 = ROIAlign((7, 7), 1 / 16.0, 2)(feat.to(DEV), g["rois"].to(DEV)).cpu()
    torch.testing.assert_close(out2, g["out_sr2"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("channels,n_rois", [(1024, 64), (132, 17)])
def test_roi_align_matches_oracle(channels, n_rois):
    from ait_b200 import ops, synth
    from ait_b200.roi_layers import ROIAlign
    feat = torch.stack([synth.c4_map(u, channels=channels) for u in range(2)])
    rois = torch.cat([synth.random_rois(u, n_rois, batch_index=u) for u in range(2)])
    ref = torch.from_numpy(c_ops.roi_align_forward(feat.numpy(), rois.numpy(), 1 / 16.0, 7, 7, 0))
    out = ROIAlign((7, 7), 1 / 16.0, 0)(feat.to(DEV), rois.to(DEV))
    torch.testing.assert_close(out.cpu(), ref, rtol=1e-5, atol=1e-6)
    # token-major layout (what feeds the enc_emb GEMM) is the same numbers transposed
    nhwc = feat.permute(0, 2, 3, 1).contiguous().to(DEV)
    tok = ops.roi_align_forward(nhwc, rois.to(DEV), 1 / 16.0, 7, 7, 0, token_major=True)
    assert torch.equal(tok.permute(0, 2, 1).reshape(out.shape), out)
    assert ROIAlign((7, 7), 1 / 16.0, 0)(feat.to(DEV), rois[:0].to(DEV)).shape == (0, channels, 7, 7)


def test_roi_align_backward_matches_oracle():
    from ait_b200.roi_layers import roi_align
    g = torch.Generator().manual_seed(9)
    feat = torch.randn(2, 48, 38, 63, generator=g)
    from ait_b200 import synth
    rois = torch.cat([synth.random_rois(u, 12, batch_index=u) for u in range(2)])
    grad = torch.randn(24, 48, 7, 7, generator=g)
    x = feat.to(DEV).requires_grad_(True)
    y = roi_align(x, rois.to(DEV), (7, 7), 1 / 16.0, 0)
    y.backward(grad.to(DEV))
    ref = c_ops.roi_align_backward(grad.numpy(), rois.numpy(), 1 / 16.0, 7, 7, 2, 48, 38, 63, 0)
    torch.testing.assert_close(x.grad.cpu().double(), torch.from_numpy(ref), rtol=1e-4, atol=1e-5)
