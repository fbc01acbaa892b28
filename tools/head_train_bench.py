"""Time the WHOLE-HEAD training step (BASELINE config 4): ROIAlign -> AIT -> SKNet -> layer4 (pairs + queries) -> heads ->
the three RCNN losses, forward + backward (DetectionHead.training_losses), batch 16 units x 128 proposals, dropout 0,
fp32 storage / tf32 math.  CUDA events on the current stream; prints one JSON line with the per-stage split.

    python tools/head_train_bench.py [B] [P] [steps]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ait_b200 import ops, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
P = int(sys.argv[2]) if len(sys.argv) > 2 else 128
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dev = "cuda:0"
head = synth.make_head(seed=0, calibrated=True, randomize_bn=True)
for mod in head.modules():
    if hasattr(mod, "p_dropout"):
        mod.p_dropout = 0.0
head = head.to(dev).train()
maps = torch.stack([synth.c4_map(u) for u in range(B)]).to(dev).requires_grad_()
qrys = torch.stack([synth.query_feat(u) for u in range(B)]).to(dev).requires_grad_()
rois = torch.stack([synth.random_rois(u, P, batch_index=u) for u in range(B)]).to(dev)
g = torch.Generator().manual_seed(3)
label = (torch.rand(B * P, generator=g) < 0.25).long().to(dev)
tgt = (0.3 * torch.randn(B * P, 4, generator=g)).to(dev)
inw = (label > 0).float().view(-1, 1).expand(-1, 4).contiguous()
outw = inw.clone()


def step():
    head.zero_grad(set_to_none=True)
    maps.grad = None
    qrys.grad = None
    losses = head.training_losses(maps, qrys, rois, label, tgt, inw, outw)
    sum(losses).backward()
    return losses


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.reset_peak_memory_stats()
ops.launch_count(reset=True)
st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
st.record()
for _ in range(steps):
    losses = step()
en.record()
torch.cuda.synchronize()
launches = ops.launch_count() // steps
ms = st.elapsed_time(en) / steps
# forward FLOPs: 1.494 GFLOP / pair + (0.214 + 0.646) GFLOP / unit; backward = 2x (dgrad + wgrad)
flops = 3 * (B * P * 1.494e9 + B * 0.860e9)
grads = [p.grad for p in head.parameters() if p.grad is not None] + [maps.grad, qrys.grad]
print(json.dumps({"workload": "whole-head training step fwd+bwd (config 4)", "B": B, "P": P, "pairs": B * P,
                  "ms_per_step": ms, "pairs_per_s": B * P / (ms * 1e-3), "tflops_tf32": flops / (ms * 1e-3) / 1e12,
                  "forward_launches_per_step": launches, "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
                  "losses": [float(x.detach()) for x in losses], "n_grads": len(grads),
                  "finite": bool(all(torch.isfinite(x).all() for x in grads))}))
