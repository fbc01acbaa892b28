"""A few SGD steps of the detection head on the device, end to end: RPN-style candidate rois + ground truth ->
ProposalTargetLayer (device sampler) -> DetectionHead.training_losses (ROIAlign, AIT, SKNet, layer4, heads, the three RCNN
losses; forward and backward in libaitb200) -> torch.optim.SGD on the head's parameters.  Prints the loss per step.

    python tools/train_demo.py [units] [steps] [lr] [momentum] [calibrated]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ait_b200 import synth  # noqa: E402
from ait_b200.targets import ProposalTargetLayer  # noqa: E402


def make_batch(B, dev, R=300):
    """Synthetic units: C4 map, query feature, R candidate rois around two ground-truth boxes per image."""
    g = torch.Generator().manual_seed(5)
    maps = torch.stack([synth.c4_map(u) for u in range(B)]).to(dev)
    qrys = torch.stack([synth.query_feat(u) for u in range(B)]).to(dev)
    gt = torch.zeros(B, 20, 5)
    for b in range(B):
        for k in range(2):
            x1, y1 = 100 + 400 * k + 40 * torch.rand(1, generator=g), 100 + 150 * k + 40 * torch.rand(1, generator=g)
            gt[b, k] = torch.tensor([float(x1), float(y1), float(x1) + 250, float(y1) + 200, 1.0])
    rois = torch.zeros(B, R, 5)
    for b in range(B):
        base = gt[b, torch.randint(0, 2, (R,), generator=g), :4]
        jit = torch.randn(R, 4, generator=g) * torch.tensor([60.0, 50.0, 60.0, 50.0])
        far = torch.rand(R, generator=g) < 0.5                      # half of the candidates are random background boxes
        rnd = synth.random_rois(b, R)[:, 1:]
        box = torch.where(far.view(-1, 1), rnd, base + jit)
        x1, y1 = box[:, 0].clamp(0, 940), box[:, 1].clamp(0, 540)
        rois[b, :, 0] = b
        rois[b, :, 1:] = torch.stack([x1, y1, torch.maximum(box[:, 2], x1 + 16).clamp(max=999),
                                      torch.maximum(box[:, 3], y1 + 16).clamp(max=599)], 1)
    return maps, qrys, rois.to(dev), gt.to(dev), torch.full((B,), 2, dtype=torch.long, device=dev)


def run(B=2, steps=8, lr=1e-2, dev="cuda:0", verbose=True, mom=0.0, calibrated=False):
    head = synth.make_head(seed=0, calibrated=calibrated, randomize_bn=True)   # stock init: the reference's normal_init
    from ait_b200.system.Models import set_dropout
    set_dropout(head, 0.0, 0.0)
    head = head.to(dev).train()
    for n, p in head.named_parameters():           # frozen BatchNorm (set_bn_fix), like the reference
        if ".bn" in n or "downsample.1" in n:
            p.requires_grad_(False)
    sampler = ProposalTargetLayer(2, rng="device", seed=3)
    maps, qrys, all_rois, gt, nb = make_batch(B, dev)
    rois, label, tgt, inw, outw = sampler(all_rois, gt, nb)          # one fixed sample: the loss must go down on it
    label = label.view(-1).long()
    opt = torch.optim.SGD([p for p in head.parameters() if p.requires_grad], lr=lr, momentum=mom)
    hist = []
    for it in range(steps):
        opt.zero_grad(set_to_none=True)
        losses = head.training_losses(maps, qrys, rois, label, tgt, inw, outw)
        total = sum(losses)
        total.backward()
        opt.step()
        hist.append([float(x.detach()) for x in losses])
        if verbose:
            print("step %d  cls %.4f  margin %.4f  bbox %.4f  total %.4f  (fg rois: %d of %d)"
                  % (it, *hist[-1], sum(hist[-1]), int((label > 0).sum()), label.numel()))
    return hist


if __name__ == "__main__":
    a = sys.argv
    run(int(a[1]) if len(a) > 1 else 2, int(a[2]) if len(a) > 2 else 8, float(a[3]) if len(a) > 3 else 1e-2,
        mom=float(a[4]) if len(a) > 4 else 0.0, calibrated=len(a) > 5 and a[5] == "calibrated")
