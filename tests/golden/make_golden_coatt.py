"""Golden outputs of the reference's co-attention block (`B.CoAttention` as built by the detector) on seeded inputs.

    python tests/golden/make_golden_coatt.py     (build container only: needs /root/reference)
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import ref_import  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def inputs(seed=29, B=2, H=19, W=31):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, 1024, H, W, generator=g).relu(), torch.randn(B, 1024, 8, 8, generator=g).relu()


def weights(seed=29):
    """state_dict with non-trivial GroupNorm parameters (the reference initialises them to 0 = identity block)."""
    g = torch.Generator().manual_seed(seed + 1)
    sd = {}
    for name in ("emb", "rho", "phi"):
        sd[name + ".weight"] = torch.randn(512, 1024, 1, 1, generator=g) * 0.03
        sd[name + ".bias"] = torch.randn(512, generator=g) * 0.1
    for name in ("omega", "theta"):
        sd[name + ".0.weight"] = torch.randn(1024, 512, 1, 1, generator=g) * 0.05
        sd[name + ".0.bias"] = torch.randn(1024, generator=g) * 0.1
        sd[name + ".1.weight"] = torch.rand(1024, generator=g) + 0.5
        sd[name + ".1.bias"] = torch.randn(1024, generator=g) * 0.2
    return sd


def main():
    torch.set_num_threads(8)
    ref_import.install()
    from model.modules import blocks_coatt_transformer_sk as Bk
    m = Bk.CoAttention(in_ch=1024, c_hidden=512, with_residual=True, normlization="division").eval()
    m.load_state_dict(weights(), strict=True)
    x_img, x_qry = inputs()
    with torch.no_grad():
        non_img, non_qry = m(x_img, x_qry)
    torch.save(dict(seed=29, non_img_s=non_img[:, ::32].clone(), non_qry_s=non_qry[:, ::16].clone(),
                    keys=sorted(m.state_dict().keys())), os.path.join(OUT, "coattention.pt"))
    print("wrote coattention.pt; |non_img - x_img| max", float((non_img - x_img).abs().max()),
          "|non_qry - x_qry| max", float((non_qry - x_qry).abs().max()))


if __name__ == "__main__":
    main()
