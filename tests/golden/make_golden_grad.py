"""Golden GRADIENTS of the AIT module (BASELINE config 4) from the UNMODIFIED reference
(model.system.Models.Transformer, .train(), every nn.Dropout set to p = 0) run on CPU fp32 with torch autograd.

    python tests/golden/make_golden_grad.py     (build container only: needs /root/reference)

Writes tests/golden/ait_grad.pt: the output, the input gradients and, for every one of the 46
parameters, the gradient's L2 norm plus a strided sample of it (full tensor when small).
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from ait_b200 import synth  # noqa: E402
from oracle import ref_import  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SAMPLE = 2048


def inputs():
    g = torch.Generator().manual_seed(13)
    xp = torch.rand(2, 1024, 7, 7, generator=g)
    xq = torch.rand(1, 1024, 8, 8, generator=g)
    gout = torch.randn(2, 1024, 8, 8, generator=g)
    return xp, xq, gout


def sample(t):
    f = t.reshape(-1)
    if f.numel() <= SAMPLE:
        return f.clone(), 1
    stride = f.numel() // SAMPLE
    return f[::stride][:SAMPLE].clone(), stride


def main():
    torch.set_num_threads(8)
    ref_import.install()
    head = synth.make_head(seed=0, calibrated=True, randomize_bn=True)
    T = ref_import.ref_transformer(dropout=0.0).train()
    # ScaledDotProductAttention keeps attn_dropout = 0.1 whatever the constructor says (system/Modules.py:9-14,
    # SubLayers.py:55 does not forward `dropout`): switch every nn.Dropout off for a deterministic graph
    for mod in T.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    T.load_state_dict(head.transformer.state_dict(), strict=True)
    xp, xq, gout = inputs()
    xp.requires_grad_()
    xq.requires_grad_()
    out = T(x_props=xp, x_query=xq)
    out.backward(gout)
    params = {}
    for name, p in T.named_parameters():
        if p.grad is None:
            continue
        s, stride = sample(p.grad)
        params[name] = dict(norm=float(p.grad.double().norm()), sample=s, stride=stride)
    torch.save(dict(seed=13, out_s=out.detach()[:, ::8].clone(), grad_props_s=xp.grad[:, ::4].clone(),
                    grad_props_norm=float(xp.grad.double().norm()), grad_query_s=xq.grad[:, ::4].clone(),
                    grad_query_norm=float(xq.grad.double().norm()), params=params),
               os.path.join(OUT, "ait_grad.pt"))
    print("wrote ait_grad.pt:", len(params), "parameter gradients; |grad_props| = %.4f" % float(xp.grad.norm()))


if __name__ == "__main__":
    main()
