"""Time the AIT training step (BASELINE config 4: forward + backward of Transformer.forward, batch 16 x 128
proposals, dropout 0) with CUDA events and print one JSON line.

    python tools/train_step_bench.py [B] [P] [steps]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ait_b200.system.Models import Transformer  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
P = int(sys.argv[2]) if len(sys.argv) > 2 else 128
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dev = "cuda:0"
torch.manual_seed(0)
m = Transformer(n_layers=1, dropout=0.0, n_position=64, attn_dropout=0.0).to(dev).train()
g = torch.Generator(device=dev).manual_seed(1)
xp = torch.randn(B * P, 1024, 7, 7, device=dev, generator=g).relu().requires_grad_()
xq = torch.randn(B, 1024, 8, 8, device=dev, generator=g).relu().requires_grad_()
gout = torch.randn(B * P, 1024, 8, 8, device=dev, generator=g)


def step():
    m.zero_grad(set_to_none=True)
    out = m(xp, xq)
    out.backward(gout)
    return out


for _ in range(2):
    step()
torch.cuda.synchronize()
st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
st.record()
for _ in range(steps):
    out = step()
en.record()
torch.cuda.synchronize()
ms = st.elapsed_time(en) / steps
# forward 848.4 MFLOP / pair (+ 214 MFLOP per unit), backward = 2x (dgrad + wgrad)
flops = 3 * (B * P * 848.4e6 + B * 214.0e6)
print(json.dumps({"workload": "AIT training step fwd+bwd (config 4)", "B": B, "P": P, "pairs": B * P,
                  "ms_per_step": ms, "pairs_per_s": B * P / (ms * 1e-3), "tflops_tf32": flops / (ms * 1e-3) / 1e12,
                  "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
                  "finite": bool(torch.isfinite(out).all() and torch.isfinite(xp.grad).all())}))
