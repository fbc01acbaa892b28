"""Time the WHOLE-HEAD training step (BASELINE config 4): ROIAlign -> AIT -> SKNet -> layer4 (pairs + queries) -> heads ->
the three RCNN losses, forward + backward (DetectionHead.training_losses), batch 16 units x 128 proposals, dropout 0,
fp32 storage / tf32 math.  CUDA events on the current stream; prints one JSON line with the per-stage split.

    python tools/head_train_bench.py [B] [P] [steps] [--graph]

--graph: additionally capture one whole step (forward + backward) into a CUDA graph and time its replay (the step has
no device->host read and the library never allocates or synchronises, so it is capturable as is).
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ait_b200 import ops, synth  # noqa: E402

want_graph = "--graph" in sys.argv
argv = [a for a in sys.argv if a != "--graph"]
B = int(argv[1]) if len(argv) > 1 else 16
P = int(argv[2]) if len(argv) > 2 else 128
steps = int(argv[3]) if len(argv) > 3 else 5
dev = "cuda:0"
head = synth.make_head(seed=0, calibrated=True, randomize_bn=True)
from ait_b200.system.Models import set_dropout
set_dropout(head, 0.0, 0.0)
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
eager_losses = [float(x.detach()) for x in losses]
del losses            # no reference to the eager autograd graph (its AccumulateGrad nodes live on the default stream)
graph_ms = None
if want_graph:
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        cg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(cg):
            g_losses = step()
        for _ in range(2):
            cg.replay()
        torch.cuda.synchronize()
        st.record()
        for _ in range(steps):
            cg.replay()
        en.record()
        torch.cuda.synchronize()
        graph_ms = st.elapsed_time(en) / steps
        same = all(abs(float(a.detach()) - b) < 1e-5 * max(1.0, abs(b)) for a, b in zip(g_losses, eager_losses))
        graph_ms = {"ms_per_step": graph_ms, "pairs_per_s": B * P / (graph_ms * 1e-3), "losses_match_eager": bool(same)}
    except Exception as e:  # report, do not hide
        graph_ms = {"failed": "%s: %s" % (type(e).__name__, str(e)[:300])}
print(json.dumps({"cuda_graph": graph_ms, "workload": "whole-head training step fwd+bwd (config 4)", "B": B, "P": P, "pairs": B * P,
                  "ms_per_step": ms, "pairs_per_s": B * P / (ms * 1e-3), "tflops_tf32": flops / (ms * 1e-3) / 1e12,
                  "forward_launches_per_step": launches, "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
                  "losses": eager_losses, "n_grads": len(grads),
                  "finite": bool(all(torch.isfinite(x).all() for x in grads))}))
