"""Golden vectors for the training step WITH dropout: the UNMODIFIED reference Transformer
(model.system.Models.Transformer, .train(), dropout = 0.1 and the hard-wired attention dropout 0.1) on CPU fp32 with torch
autograd, its ten nn.Dropout instances fed the seeded masks of oracle/drop_masks.py instead of drawing their own (the
reference's draws come from torch's generator state and cannot be matched by any other implementation).  The encoder
embedding mask covers all 64 rows, zero-padded rows included, exactly like nn.Dropout does (those rows are queries of the
encoder self-attention and enter its selective-head gate, so their masks matter).

    python tests/golden/make_golden_drop.py       (build container only: needs /root/reference)
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ait_b200 import synth  # noqa: E402
from oracle import drop_masks, ref_import  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SAMPLE = 4096
BS, P, SEED, MASK_SEED = 2, 2, 17, 1234


def inputs():
    g = torch.Generator().manual_seed(SEED)
    xp = torch.rand(BS * P, 1024, 7, 7, generator=g)
    xq = torch.rand(BS, 1024, 8, 8, generator=g)
    gout = torch.randn(BS * P, 1024, 8, 8, generator=g)
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
    T = ref_import.ref_transformer(dropout=0.1).train()
    T.load_state_dict(head.transformer.state_dict(), strict=True)
    masks = drop_masks.make_masks(MASK_SEED, BS, P, 0.1, 0.1)
    drop_masks.inject_into_reference(T, masks, P)
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
    torch.save(dict(seed=SEED, mask_seed=MASK_SEED, bs=BS, num_props=P, p=0.1, p_attn=0.1,
                    out_s=out.detach()[:, ::8].clone(), grad_props_s=xp.grad[:, ::4].clone(),
                    grad_query_s=xq.grad[:, ::4].clone(), params=params),
               os.path.join(OUT, "ait_drop.pt"))
    print("wrote ait_drop.pt:", len(params), "parameter gradients; |out| = %.4f" % float(out.norm()))


if __name__ == "__main__":
    main()
