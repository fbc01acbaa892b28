"""Adaptive Image Transformer -- parameter-compatible mirror of the reference's
`model.system.Models.Transformer` (lib/model/system/Models.py:174-280, identical to
lib/model/transformer/Models.py at the shipped hyper-parameters), executed by libaitb200.

The nn.Module tree below exists to (1) own the parameters/buffers under EXACTLY the reference's
`state_dict` keys (48 keys; reference checkpoints load with strict=True) and (2) expose the
reference's constructor and `forward(x_props, x_query)` signature.  No layer has a torch
forward of its own: `Transformer.forward` packs the weights once and calls `aitb_ait_forward`
(tcgen05 GEMMs with fused bias / positional / residual / LayerNorm epilogues + the selective-head
attention kernel).  There is no PyTorch fallback.
"""
import numpy as np
import torch
import torch.nn as nn

from .. import packing


def _sinusoid_table(n_position, d_hid):
    # same float64 construction as PositionalEncoding._get_sinusoid_encoding_table (Models.py:32-45)
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)[None, :]
    angle = pos / np.power(10000, 2 * (j // 2) / d_hid)
    table = np.empty_like(angle)
    table[:, 0::2] = np.sin(angle[:, 0::2])
    table[:, 1::2] = np.cos(angle[:, 1::2])
    return torch.FloatTensor(table).unsqueeze(0)


class _Holder(nn.Module):
    """Parameter container: these sub-modules are never called on their own."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("%s is executed by the fused libaitb200 engine; call Transformer.forward"
                           % type(self).__name__)


class PositionalEncoding(_Holder):
    def __init__(self, d_hid, n_position=200):
        super().__init__()
        self.register_buffer("pos_table", _sinusoid_table(n_position, d_hid))


class SHBlock(_Holder):
    """Selective-head gate (SubLayers.py:9-39): only `sk` carries parameters."""

    def __init__(self, n_head, d_v):
        super().__init__()
        self.sk = nn.Linear(d_v, d_v * n_head)


class MultiHeadAttention(_Holder):
    def __init__(self, n_head, d_model, d_k, d_v, dropout=0.1):
        super().__init__()
        self.n_head, self.d_k, self.d_v = n_head, d_k, d_v
        self.w_qs = nn.Linear(d_model, n_head * d_k, bias=False)
        self.w_ks = nn.Linear(d_model, n_head * d_k, bias=False)
        self.w_vs = nn.Linear(d_model, n_head * d_v, bias=False)
        self.sh = SHBlock(n_head=n_head, d_v=d_v)
        self.fc = nn.Linear(d_v, d_model, bias=False)
        self.layer_norm = nn.LayerNorm(d_model, eps=1e-6)
        self.p_dropout = dropout
        # ScaledDotProductAttention(temperature, attn_dropout=0.1): SubLayers.py:55 never forwards `dropout`, so the
        # attention-probability dropout is 0.1 in .train() whatever the constructor says (Modules.py:9-14,24)
        self.p_attn_dropout = 0.1


class PositionwiseFeedForward(_Holder):
    def __init__(self, d_in, d_hid, dropout=0.1):
        super().__init__()
        self.w_1 = nn.Linear(d_in, d_hid)
        self.w_2 = nn.Linear(d_hid, d_in)
        self.layer_norm = nn.LayerNorm(d_in, eps=1e-6)
        self.p_dropout = dropout


class EncoderLayer(_Holder):
    def __init__(self, d_model, d_inner, n_head, d_k, d_v, dropout=0.1):
        super().__init__()
        self.slf_attn = MultiHeadAttention(n_head, d_model, d_k, d_v, dropout=dropout)
        self.pos_ffn = PositionwiseFeedForward(d_model, d_inner, dropout=dropout)


class DecoderLayer(_Holder):
    def __init__(self, d_model, d_inner, n_head, d_k, d_v, dropout=0.1):
        super().__init__()
        self.slf_attn = MultiHeadAttention(n_head, d_model, d_k, d_v, dropout=dropout)
        self.enc_attn = MultiHeadAttention(n_head, d_model, d_k, d_v, dropout=dropout)
        self.pos_ffn = PositionwiseFeedForward(d_model, d_inner, dropout=dropout)


class Encoder(_Holder):
    def __init__(self, d_word_vec, n_layers, n_head, d_k, d_v, d_model, d_inner, pad_idx, dropout=0.1,
                 n_position=200):
        super().__init__()
        self.position_enc = PositionalEncoding(d_word_vec, n_position=n_position)
        self.layer_stack = nn.ModuleList(
            [EncoderLayer(d_model, d_inner, n_head, d_k, d_v, dropout=dropout) for _ in range(n_layers)])
        self.layer_norm = nn.LayerNorm(d_model, eps=1e-6)


class Decoder(_Holder):
    def __init__(self, d_word_vec, n_layers, n_head, d_k, d_v, d_model, d_inner, pad_idx, n_position=200,
                 dropout=0.1):
        super().__init__()
        self.position_enc = PositionalEncoding(d_word_vec, n_position=n_position)
        self.layer_stack = nn.ModuleList(
            [DecoderLayer(d_model, d_inner, n_head, d_k, d_v, dropout=dropout) for _ in range(n_layers)])
        self.layer_norm = nn.LayerNorm(d_model, eps=1e-6)


_keep_saved_for_tests = [False, None]     # [enabled, the last training forward's saved-activation buffer]


def set_dropout(module, p=None, p_attn=None):
    """Set the training-mode dropout probabilities of every AIT sub-layer below `module` (a Transformer or anything that
    contains one): p -> the nn.Dropout sites, p_attn -> the attention probabilities (the reference hard-wires 0.1 there).
    set_dropout(m, 0.0, 0.0) gives the deterministic training graph (BASELINE configs[3] numbers, the gradient goldens)."""
    for mod in module.modules():
        if p is not None and hasattr(mod, "p_dropout"):
            mod.p_dropout = float(p)
        if p_attn is not None and hasattr(mod, "p_attn_dropout"):
            mod.p_attn_dropout = float(p_attn)
    return module


class _AITTrainFunction(torch.autograd.Function):
    """Transformer.forward with a hand-written backward (libaitb200: tcgen05 dgrad / wgrad GEMMs, LayerNorm /
    attention / selective-head-gate backward kernels).  The reference relies on torch autograd over
    system/Models.py:231-280; gradients match it to tf32 accuracy, or to bf16 accuracy when the module was built with
    compute_dtype=torch.bfloat16 (tests/test_gpu_train.py).  Inputs, outputs and all gradients are fp32 tensors either way."""

    @staticmethod
    def forward(ctx, x_props, x_query, module, tm_out, drop, *params):
        # weights change every step: repack.  compute_dtype bfloat16 -> the bf16 training configuration, else fp32 storage / tf32 math
        engine = packing.HeadEngine(transformer=module, dtype="bf16" if packing.L.mode_name(module.compute_dtype) == "bf16" else "tf32")
        engine.set_train_dropout(*drop)             # (p_drop, p_attn, seed): the backward regenerates the same masks
        out, saved = engine.ait_forward_train(x_props, x_query, token_major_out=tm_out)
        ctx.engine, ctx.saved, ctx.tm_out = engine, saved, tm_out
        if _keep_saved_for_tests[0]:
            _keep_saved_for_tests[1] = saved       # the parity tests read the device's FFN ReLU decisions out of it
        ctx.bs, ctx.num_props = x_query.shape[0], x_props.shape[0] // x_query.shape[0]
        ctx.in_dtypes = (x_props.dtype, x_query.dtype)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        if ctx.saved is None:      # the activation buffer (GBs) is released after the first backward, like autograd's own buffers
            raise RuntimeError("ait_b200.Transformer: trying to backward through the AIT training step a second time: its "
                               "saved activations were freed after the first backward (retain_graph=True is not supported; "
                               "run the forward again)")
        g_props, g_query, g_params = ctx.engine.ait_backward(grad_out, ctx.saved, ctx.bs, ctx.num_props,
                                                             token_major_grad=ctx.tm_out)
        ctx.saved = None
        return (g_props.to(ctx.in_dtypes[0]), g_query.to(ctx.in_dtypes[1]), None, None, None) + tuple(g_params)


class Transformer(nn.Module):
    """Same constructor as the reference (Models.py:177-181).  Supported envelope of the fused engine:
    d_model = d_word_vec = 512, d_inner = 2048, n_layers = 1, n_head = 8, d_k = d_v = 64,
    n_position >= 64 -- what `_fasterRCNN` and adaptive_image_transformer.py instantiate.  Anything
    else raises (no silent fallback)."""

    def __init__(self, src_pad_idx=1, trg_pad_idx=1, d_word_vec=512, d_model=512, d_inner=2048, n_layers=6,
                 n_head=8, d_k=64, d_v=64, dropout=0.1, n_position=200, trg_emb_prj_weight_sharing=True,
                 emb_src_trg_weight_sharing=True, compute_dtype=torch.float32, attn_dropout=None):
        super().__init__()
        if d_model != d_word_vec:
            raise AssertionError("To facilitate the residual connections, the dimensions of all module "
                                 "outputs shall be the same.")
        if (d_model, d_inner, n_layers, n_head, d_k, d_v) != (512, 2048, 1, 8, 64, 64) or n_position < 64:
            raise NotImplementedError(
                "ait_b200.Transformer supports the shipped AIT configuration only "
                "(d_model=512, d_inner=2048, n_layers=1, n_head=8, d_k=d_v=64, n_position>=64); got "
                "d_model=%d d_inner=%d n_layers=%d n_head=%d d_k=%d d_v=%d n_position=%d"
                % (d_model, d_inner, n_layers, n_head, d_k, d_v, n_position))
        self.src_pad_idx, self.trg_pad_idx = src_pad_idx, trg_pad_idx
        self.channels = d_word_vec
        self.compute_dtype = compute_dtype
        self.enc_emb = nn.Sequential(nn.Conv2d(d_word_vec * 2, d_word_vec, kernel_size=1, bias=True))
        self.dec_emb = nn.Sequential(nn.Conv2d(d_word_vec * 2, d_word_vec, kernel_size=1, bias=True))
        self.encoder = Encoder(n_position=n_position, d_word_vec=d_word_vec, d_model=d_model, d_inner=d_inner,
                               n_layers=n_layers, n_head=n_head, d_k=d_k, d_v=d_v, pad_idx=src_pad_idx,
                               dropout=dropout)
        self.decoder = Decoder(n_position=n_position, d_word_vec=d_word_vec, d_model=d_model, d_inner=d_inner,
                               n_layers=n_layers, n_head=n_head, d_k=d_k, d_v=d_v, pad_idx=trg_pad_idx,
                               dropout=dropout)
        self.dec_trans = nn.Sequential(nn.Conv2d(d_word_vec, d_word_vec * 2, kernel_size=1, bias=True))
        for p in self.parameters():  # Models.py:214-216
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        self._engine = None
        self.last_dropout_seed = 0
        if attn_dropout is not None:       # not a reference argument: the reference hard-wires 0.1 (Modules.py:9)
            for m in self.modules():
                if hasattr(m, "p_attn_dropout"):
                    m.p_attn_dropout = float(attn_dropout)

    def _draw_dropout(self):
        """(p_drop, p_attn, seed) of one training step.  The engine has ONE probability per kind: every holder must agree."""
        ps = {float(m.p_dropout) for m in self.modules() if hasattr(m, "p_dropout")}
        pa = {float(m.p_attn_dropout) for m in self.modules() if hasattr(m, "p_attn_dropout")}
        if len(ps) != 1 or len(pa) != 1:
            raise RuntimeError("ait_b200.Transformer: all sub-layers must share one dropout / one attention-dropout "
                               "probability (got %s / %s)" % (sorted(ps), sorted(pa)))
        p, p_attn = ps.pop(), pa.pop()
        if not (0.0 <= p < 1.0 and 0.0 <= p_attn < 1.0):
            raise RuntimeError("ait_b200.Transformer: dropout probabilities must be in [0, 1)")
        seed = 0
        if p > 0.0 or p_attn > 0.0:
            seed = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())
        self.last_dropout_seed = seed
        return (p, p_attn, seed)

    def invalidate(self):
        """Call after changing parameters in place (load_state_dict does it automatically)."""
        self._engine = None

    def _load_from_state_dict(self, *a, **k):
        super()._load_from_state_dict(*a, **k)
        self._engine = None

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def forward(self, x_props, x_query, token_major_out=False):
        """x_props [bs*num_props, 1024, 7, 7], x_query [bs, 1024, 8, 8] -> [bs*num_props, 1024, 8, 8]
        (token_major_out, training step only: the [bs*num_props, 64, 1024] token-major, tf32-rounded result without the NCHW
        copy, for `sk_train.sknet_train(..., channels_last_in=True)`; its gradient comes back in the same layout)
        (Models.py:231-280).  .eval(): inference engine (no autograd graph, dropout = identity, like the reference in
        .eval()).  .train(): differentiable training step with the library's own backward; dropout as in the reference's
        .train(): p = `dropout` at the seven nn.Dropout sites (Models.py:98,152; SubLayers.py:97,182) and 0.1 on the attention
        probabilities (Modules.py:24; `attn_dropout=` overrides it, 0.0 for a deterministic graph).  The masks come from
        a counter-based generator keyed by a per-step seed drawn from torch's CPU generator (torch.manual_seed makes a
        run reproducible; `last_dropout_seed` records it); the decoder-side sites draw ONE mask per unit, shared by the
        unit's proposals, because the query side runs once per unit (the reference recomputes it, with fresh masks, per
        proposal)."""
        if x_props.dim() != 4 or x_query.dim() != 4:
            raise RuntimeError("expected x_props [bp,1024,7,7] and x_query [bs,1024,8,8]")
        bp, c_p, h_p, w_p = x_props.shape
        bs, c_q, h_q, w_q = x_query.shape
        if (c_p, h_p, w_p) != (2 * self.channels, 7, 7) or (c_q, h_q, w_q) != (2 * self.channels, 8, 8):
            raise RuntimeError("ait_b200.Transformer: supported shapes are x_props [bp,1024,7,7] and x_query "
                               "[bs,1024,8,8]; got %s and %s" % (tuple(x_props.shape), tuple(x_query.shape)))
        if bs == 0 or bp == 0 or bp % bs != 0:
            raise RuntimeError("bp=%d must be a positive multiple of bs=%d (num_props = bp // bs)" % (bp, bs))
        if self.training and torch.is_grad_enabled() and (
                x_props.requires_grad or x_query.requires_grad or any(p.requires_grad for p in self.parameters())):
            # training step (BASELINE config 4): forward keeping activations + hand-written backward, fp32 storage with
            # tf32 tensor-core math
            sd = dict(self.named_parameters())
            params = [sd[n] for n in packing.HeadEngine.ait_param_names()]
            return _AITTrainFunction.apply(x_props, x_query, self, bool(token_major_out), self._draw_dropout(), *params)
        if token_major_out:
            raise RuntimeError("ait_b200.Transformer: token_major_out is a training-step hand-over (module in .train(), grad enabled)")
        fp = packing.fingerprint(self)          # in-place parameter updates since the last packing?
        if self._engine is None or self._engine_fp != fp:
            self._engine_fp = fp
            self._engine = packing.HeadEngine(transformer=self, dtype=self.compute_dtype)
        return self._engine.ait_forward(x_props, x_query)
