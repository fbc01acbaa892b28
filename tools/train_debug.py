"""Per-tensor gradient errors of the AIT training step against fp64 autograd of the CPU oracle."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ait_b200.system.Models import Transformer  # noqa: E402
from oracle import head_oracle  # noqa: E402

B, P = 2, 3
DEV = "cuda:0"
torch.manual_seed(0)
m = Transformer(n_layers=1, dropout=0.0, n_position=64, attn_dropout=0.0).train()
if len(sys.argv) > 1 and sys.argv[1] == "smooth":     # every ReLU active: gradients are smooth in the weights
    with torch.no_grad():
        m.encoder.layer_stack[0].pos_ffn.w_1.bias += 5.0
        m.decoder.layer_stack[0].pos_ffn.w_1.bias += 5.0
g = torch.Generator().manual_seed(11)
x_props = torch.randn(B * P, 1024, 7, 7, generator=g).relu()
x_query = torch.randn(B, 1024, 8, 8, generator=g).relu()
gout = torch.randn(B * P, 1024, 8, 8, generator=g)
sd = {k: v.detach().double().clone().requires_grad_(v.is_floating_point() and "pos_table" not in k)
      for k, v in m.state_dict().items()}
xp, xq = x_props.double().requires_grad_(), x_query.double().requires_grad_()
ref = head_oracle.ait_forward(sd, xp, xq, dtype=torch.float64)
ref.backward(gout.double())
m = m.to(DEV)
xp2, xq2 = x_props.to(DEV).requires_grad_(), x_query.to(DEV).requires_grad_()
out = m(xp2, xq2)
out.backward(gout.to(DEV))
torch.cuda.synchronize()


def rel(a, b):
    a = a.detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max()), float((a - b).norm() / b.norm())


print("out", rel(out, ref.detach()))
print("grad x_props", rel(xp2.grad, xp.grad))
print("grad x_query", rel(xq2.grad, xq.grad))
for name, p in m.named_parameters():
    print("%-55s max-rel %.2e  l2-rel %.2e" % ((name,) + rel(p.grad, sd[name].grad)))
