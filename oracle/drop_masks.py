"""TEST INFRASTRUCTURE ONLY.  Seeded dropout multipliers (0 | 1 / (1 - p)) in the site layout of
`ait_b200.packing.HeadEngine.dropout_masks` / `oracle.head_oracle.ait_forward(drop=...)`, and their injection into the
UNMODIFIED reference Transformer (every nn.Dropout INSTANCE of lib/model/system gets a forward that multiplies by the
given mask -- no reference code is changed).  Used by tests/golden/make_golden_drop.py and tests/test_oracle_pins.py to
pin the oracle's dropout sites to the reference's own nn.Dropout call sites (Models.py:98,152; SubLayers.py:97,182;
Modules.py:24)."""
import torch

ROW_SITES = {"enc_emb": False, "dec_emb": True, "enc_slf_fc": False, "dec_slf_fc": True, "dec_enc_fc": False,
             "enc_ffn": False, "dec_ffn": False}             # name -> one mask per UNIT (shared by its proposals)?
ATTN_SITES = {"enc_slf_attn": False, "dec_slf_attn": True, "dec_enc_attn": False}


def make_masks(seed, bs, num_props, p=0.1, p_attn=0.1):
    g = torch.Generator().manual_seed(seed)
    bp = bs * num_props
    out = {}
    for name, per_unit in ROW_SITES.items():
        n = bs if per_unit else bp
        out[name] = (torch.rand(n * 64, 512, generator=g) >= p).float() / (1.0 - p)
    for name, per_unit in ATTN_SITES.items():
        n = bs if per_unit else bp
        out[name] = (torch.rand(n, 8, 64, 64, generator=g) >= p_attn).float() / (1.0 - p_attn)
    return out


def inject_into_reference(T, masks, num_props):
    """T: the reference's model.system.Models.Transformer in .train().  Each nn.Dropout instance multiplies by its site's
    mask instead of drawing one; the per-unit decoder-side masks are repeated over the unit's proposals, which is how
    the reference lays out its per-proposal copies of the query (Models.py:250-253: unit-major)."""
    def rows(name):
        m = masks[name].view(-1, 64, 512)
        return m.repeat_interleave(num_props, dim=0) if ROW_SITES[name] else m

    def heads(name):
        m = masks[name]
        return m.repeat_interleave(num_props, dim=0) if ATTN_SITES[name] else m

    def patch(drop_module, mask):
        drop_module.forward = lambda x, _m=mask: x * _m.to(x.dtype)

    enc, dec = T.encoder.layer_stack[0], T.decoder.layer_stack[0]
    patch(T.encoder.dropout, rows("enc_emb"))
    patch(T.decoder.dropout, rows("dec_emb"))
    patch(enc.slf_attn.attention.dropout, heads("enc_slf_attn"))
    patch(enc.slf_attn.dropout, rows("enc_slf_fc"))
    patch(enc.pos_ffn.dropout, rows("enc_ffn"))
    patch(dec.slf_attn.attention.dropout, heads("dec_slf_attn"))
    patch(dec.slf_attn.dropout, rows("dec_slf_fc"))
    patch(dec.enc_attn.attention.dropout, heads("dec_enc_attn"))
    patch(dec.enc_attn.dropout, rows("dec_enc_fc"))
    patch(dec.pos_ffn.dropout, rows("dec_ffn"))
    n_drop = sum(isinstance(m, torch.nn.Dropout) for m in T.modules())
    assert n_drop == 10, "the reference Transformer has %d nn.Dropout instances, expected 10" % n_drop
    return T
