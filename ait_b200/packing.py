"""Weight packing + the host side of the head engine.

`HeadEngine` turns the parameters of the mirror modules (system/Models.py, modules.py, head.py)
into the flat, GEMM-ready buffers `aitb_head_weights` points at (include/aitb200.h):

  * every projection / convolution becomes a K-major [N, K] matrix in the compute dtype
    (fp32 pre-rounded to tf32, or bf16); w_qs/w_ks/w_vs are stacked into one [1536, 512] matrix
  * 3x3 kernels are re-ordered tap-major ([out, ky, kx, in]) to match the shifted-TMA K loop
  * frozen BatchNorm (eval mode, resnet_coatt_transformer_sk.py:429-435,451-474) is folded into
    the preceding conv: w' = w * gamma / sqrt(var + eps), b' = beta - mean * gamma / sqrt(var + eps)
  * biases, LayerNorm parameters, positional tables and the tiny score/box heads stay fp32

Packing happens once (off the hot path); the per-step call is a single `aitb_head_forward`.
"""
import ctypes as C

import torch

from . import _lib as L
from . import ops


def round_to_tf32(t):
    """fp32 -> nearest tf32 (ties away from zero), like cvt.rna.tf32.f32."""
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def fingerprint(module):
    """(storage pointer, version counter) of every parameter / buffer of a module tree.  The packed GEMM-ready copies are
    snapshots: an in-place update (optimizer.step(), param.data.copy_(), BatchNorm statistics) bumps `_version`, a
    re-assignment changes the pointer -- either makes the cached engine stale, and every cache checks this before use."""
    import itertools
    return tuple((t.data_ptr(), t._version) for t in itertools.chain(module.parameters(), module.buffers()))


class HeadEngine:
    # Precision plan of the "fp32" configuration (include/aitb200.h AITB_PLAN_*): PLAN_ENC_ONEPASS runs the five
    # encoder-side GEMMs as one tensor-core pass on fp16 hi planes.  AITB_PRECISION_PLAN=0 in the environment (read when the
    # engine is built) restores three passes everywhere -- the A/B switch of tools/ and the tests.
    DEFAULT_PLAN = L.PLAN_ENC_ONEPASS

    def __init__(self, transformer=None, sk=None, top=None, cls_score=None, bbox_pred=None,
                 dtype=torch.float32, round_acts=True, plan=None):
        import os
        L.load()
        self.mode = L.mode_name(dtype)             # "fp32" (split bf16 x3) | "tf32" | "bf16"
        if plan is None:
            plan = int(os.environ.get("AITB_PRECISION_PLAN", self.DEFAULT_PLAN))
        self.plan = int(plan) if self.mode == "fp32" else 0
        self.enc_f16 = bool(self.plan & L.PLAN_ENC_ONEPASS)      # pooled / enc_out taps and the encoder weights are fp16 planes
        self.dt = L.MODES[self.mode]
        self.dtype = L.storage_dtype(self.mode)    # element type of the activation / weight buffers
        self.split = self.mode == "fp32"
        self.planes = 2 if self.split else 1
        self.round_acts = bool(round_acts) and self.mode == "tf32"
        self._keep = []          # packed tensors stay alive as long as the engine
        self.w = L.HeadWeights()
        self.w.dtype = self.dt
        self.w.round_tf32 = 1 if self.round_acts else 0
        self.w.plan = self.plan
        self.device = None
        self.has_ait = transformer is not None
        self.has_sk = sk is not None
        self.has_top = top is not None
        self.has_heads = cls_score is not None and bbox_pred is not None
        with torch.no_grad():
            if transformer is not None:
                self._pack_transformer(transformer)
            if sk is not None:
                self._pack_sk(sk)
            if top is not None:
                self._pack_top(top)
            if self.has_heads:
                self._pack_heads(cls_score, bbox_pred)
        self._ws = None

    # ------------------------------------------------------------------ packing helpers
    def _dev(self, t):
        if not t.is_cuda:
            raise RuntimeError("ait_b200: module parameters must live on a CUDA device (call .cuda()); "
                               "there is no CPU execution path")
        if self.device is None:
            self.device = t.device
        elif t.device != self.device:
            raise RuntimeError("ait_b200: all parameters must be on the same device")
        return t

    def _mat(self, t, f16=False):
        """[N, K] matrix in the compute dtype (f16: fp16 planes -- the operand of a one-pass GEMM of the precision plan)."""
        t = self._dev(t).detach().float().contiguous()
        if self.mode == "tf32":
            t = round_to_tf32(t)
        elif self.mode == "bf16":
            t = t.to(torch.bfloat16)
        else:                                       # [N, hi K | lo K]
            t = ops.split_planes(t, f16=f16 and self.enc_f16)
        t = t.contiguous()
        self._keep.append(t)
        return t

    def _f32(self, t):
        t = self._dev(t).detach().float().contiguous()
        self._keep.append(t)
        return t

    def _linear(self, dst, w, b=None, f16=False):
        m = self._mat(w, f16=f16)
        bt = self._f32(b) if b is not None else None
        dst.w = m.data_ptr()
        dst.bias = bt.data_ptr() if bt is not None else None
        return m, bt

    def _ln(self, dst, ln):
        dst.gamma = self._f32(ln.weight).data_ptr()
        dst.beta = self._f32(ln.bias).data_ptr()

    def _mha(self, dst, m, f16_rows=None):
        """f16_rows: (first, last) rows of the stacked [w_qs; w_ks; w_vs] matrix read by a one-pass GEMM (fp16 planes)."""
        wcat = torch.cat([m.w_qs.weight, m.w_ks.weight, m.w_vs.weight], dim=0)
        if f16_rows is not None and self.split and self.enc_f16:
            a, b = f16_rows
            parts = [ops.split_planes(self._dev(wcat[:a]).detach().float()), ops.split_planes(self._dev(wcat[a:b]).detach().float(), f16=True),
                     ops.split_planes(self._dev(wcat[b:]).detach().float())]
            wqkv = torch.cat([p_ for p_ in parts if p_.shape[0] > 0], dim=0).contiguous()
            self._keep.append(wqkv)
        else:
            wqkv = self._mat(wcat)
        dst.w_qkv = wqkv.data_ptr()
        dst.w_sk = self._f32(m.sh.sk.weight).data_ptr()
        dst.b_sk = self._f32(m.sh.sk.bias).data_ptr()
        dst.w_fc = self._mat(m.fc.weight).data_ptr()
        self._ln(dst.ln, m.layer_norm)

    def _ffn(self, dst, f, f16=False):
        self._linear(dst.w1, f.w_1.weight, f.w_1.bias, f16=f16)
        self._linear(dst.w2, f.w_2.weight, f.w_2.bias, f16=f16)
        self._ln(dst.ln, f.layer_norm)

    def _pack_transformer(self, t):
        w = self.w
        self._linear(w.enc_emb, t.enc_emb[0].weight.flatten(1), t.enc_emb[0].bias, f16=True)   # one-pass GEMM (plan)
        self._linear(w.dec_emb, t.dec_emb[0].weight.flatten(1), t.dec_emb[0].bias)
        self._linear(w.dec_trans, t.dec_trans[0].weight.flatten(1), t.dec_trans[0].bias)
        enc_pos = t.encoder.position_enc.pos_table
        dec_pos = t.decoder.position_enc.pos_table
        if enc_pos.shape[1] < 64 or dec_pos.shape[1] < 64:
            raise RuntimeError("pos_table must hold at least 64 positions")
        w.enc_pos = self._f32(enc_pos[0, :64]).data_ptr()
        w.dec_pos = self._f32(dec_pos[0, :64]).data_ptr()
        self._ln(w.enc_ln, t.encoder.layer_norm)
        self._ln(w.dec_ln, t.decoder.layer_norm)
        el, dl = t.encoder.layer_stack[0], t.decoder.layer_stack[0]
        self._mha(w.enc_slf, el.slf_attn, f16_rows=(0, 1536))      # encoder QKV projection: one pass (plan)
        self._ffn(w.enc_ffn, el.pos_ffn, f16=True)                 # encoder FFN: one pass (plan)
        self._mha(w.dec_slf, dl.slf_attn)
        self._mha(w.dec_enc, dl.enc_attn, f16_rows=(512, 1536))    # cross attention: K / V rows one pass, the Q rows (query side) three
        self._ffn(w.dec_ffn, dl.pos_ffn)

    @staticmethod
    def _tap_major(wconv):
        """[out, in, kh, kw] -> [out, kh*kw*in]"""
        return wconv.permute(0, 2, 3, 1).reshape(wconv.shape[0], -1)

    def _pack_sk(self, sk):
        self._sk_mats = {}
        for name, dst, blk in (("props", self.w.sk_props, sk.sk_props), ("query", self.w.sk_query, sk.sk_query)):
            c1, c3 = blk.convs[0][0], blk.convs[1][0]
            m1, b1 = self._linear(dst.conv1x1, self._tap_major(c1.weight), c1.bias)
            m3, b3 = self._linear(dst.conv3x3, self._tap_major(c3.weight), c3.bias)
            mf = self._mat(torch.cat([self._tap_major(c3.weight), self._tap_major(c1.weight)], dim=1))
            dst.w_fused = mf.data_ptr()
            self._sk_mats[name] = (m1, b1, m3, b3, mf)

    @staticmethod
    def _fold_bn(conv, bn):
        scale = bn.weight.float() / torch.sqrt(bn.running_var.float() + bn.eps)
        w = conv.weight.float() * scale.view(-1, 1, 1, 1)
        b = bn.bias.float() - bn.running_mean.float() * scale
        return w, b

    def _pack_top(self, top):
        layer4 = top[0]
        self._top_mats = []
        for i, blk in enumerate(layer4):
            dst = self.w.top[i]
            mats = {}
            for key, conv, bn in (("conv1", blk.conv1, blk.bn1), ("conv2", blk.conv2, blk.bn2),
                                  ("conv3", blk.conv3, blk.bn3)):
                wf, bf = self._fold_bn(conv, bn)
                mats[key] = self._linear(getattr(dst, key), self._tap_major(wf), bf)
            if blk.downsample is not None:
                wf, bf = self._fold_bn(blk.downsample[0], blk.downsample[1])
                mats["down"] = self._linear(dst.down, self._tap_major(wf), bf)
            self._top_mats.append(mats)

    def _pack_heads(self, cls_score, bbox_pred):
        w = self.w
        self._head_t = dict(
            w_bbox=self._f32(bbox_pred.weight), b_bbox=self._f32(bbox_pred.bias),
            w1=self._f32(cls_score[0].weight), b1=self._f32(cls_score[0].bias),
            w2=self._f32(cls_score[1].weight), b2=self._f32(cls_score[1].bias))
        w.w_bbox, w.b_bbox = self._head_t["w_bbox"].data_ptr(), self._head_t["b_bbox"].data_ptr()
        w.w_cls1, w.b_cls1 = self._head_t["w1"].data_ptr(), self._head_t["b1"].data_ptr()
        w.w_cls2, w.b_cls2 = self._head_t["w2"].data_ptr(), self._head_t["b2"].data_ptr()

    # ------------------------------------------------------------------ workspace
    def _workspace(self, nbytes):
        if self._ws is None or self._ws.numel() < nbytes + 1024:
            self._ws = None
            self._ws = torch.empty(int(nbytes) + 1024, dtype=torch.uint8, device=self.device)
        off = (-self._ws.data_ptr()) % 1024
        return self._ws[off:]

    # ------------------------------------------------------------------ entry points
    def ait_forward(self, x_props, x_query):
        if not self.has_ait:
            raise RuntimeError("engine built without the transformer")
        lib = L.load()
        ops._need_cuda(x_props, x_query)
        x_props = x_props.contiguous().float()
        x_query = x_query.contiguous().float()
        bp, bs = x_props.shape[0], x_query.shape[0]
        P = bp // bs
        out = torch.empty((bp, 1024, 8, 8), dtype=torch.float32, device=x_props.device)
        nbytes = lib.aitb_ait_workspace_bytes(bs, P, self.dt)
        ws = self._workspace(nbytes)
        with torch.cuda.device(x_props.device):
            L.check(lib.aitb_ait_forward(C.byref(self.w), L.ptr(x_props), L.ptr(x_query), bs, P, L.ptr(out),
                                         L.ptr(ws), nbytes, L.stream_ptr()))
        return out

    # ------------------------------------------------------------------ training (config 4)
    # parameter order of the gradient tuple returned by ait_backward (names relative to the Transformer module)
    @staticmethod
    def ait_param_names():
        names = ["enc_emb.0.weight", "enc_emb.0.bias", "dec_emb.0.weight", "dec_emb.0.bias",
                 "dec_trans.0.weight", "dec_trans.0.bias",
                 "encoder.layer_norm.weight", "encoder.layer_norm.bias",
                 "decoder.layer_norm.weight", "decoder.layer_norm.bias"]
        for m in ("encoder.layer_stack.0.slf_attn", "decoder.layer_stack.0.slf_attn", "decoder.layer_stack.0.enc_attn"):
            names += [m + s for s in (".w_qs.weight", ".w_ks.weight", ".w_vs.weight", ".sh.sk.weight", ".sh.sk.bias",
                                      ".fc.weight", ".layer_norm.weight", ".layer_norm.bias")]
        for m in ("encoder.layer_stack.0.pos_ffn", "decoder.layer_stack.0.pos_ffn"):
            names += [m + s for s in (".w_1.weight", ".w_1.bias", ".w_2.weight", ".w_2.bias", ".layer_norm.weight",
                                      ".layer_norm.bias")]
        return names

    def set_train_dropout(self, p_drop=0.0, p_attn=0.0, seed=0):
        """Dropout of the training step (aitb_head_weights.p_drop / p_attn / drop_seed): forward and backward of one step
        read the same three values and regenerate the same masks; the inference entry points ignore them."""
        self.w.p_drop, self.w.p_attn, self.w.drop_seed = float(p_drop), float(p_attn), int(seed)

    def dropout_masks(self, bs, P, device):
        """The multipliers (0 | 1 / (1 - p)) of every dropout site of the current (p_drop, p_attn, seed), materialised for
        parity tests: {site name: tensor}; row-wise sites [rows, 512], attention sites [G, 8, 64, 64]."""
        lib = L.load()
        bp = bs * P
        rows = {"enc_emb": (1, bp * 64), "dec_emb": (2, bs * 64), "enc_slf_fc": (3, bp * 64), "dec_slf_fc": (4, bs * 64),
                "dec_enc_fc": (5, bp * 64), "enc_ffn": (6, bp * 64), "dec_ffn": (7, bp * 64)}
        attn = {"enc_slf_attn": (8, bp), "dec_slf_attn": (9, bs), "dec_enc_attn": (10, bp)}
        out = {}
        with torch.cuda.device(device):
            for name, (site, n) in rows.items():
                t = torch.empty((n, 512), dtype=torch.float32, device=device)
                L.check(lib.aitb_dropout_mask(self.w.p_drop, self.w.drop_seed, site, n, L.ptr(t), L.stream_ptr()))
                out[name] = t
            for name, (site, g) in attn.items():
                t = torch.empty((g, 8, 64, 64), dtype=torch.float32, device=device)
                L.check(lib.aitb_attn_dropout_mask(self.w.p_attn, self.w.drop_seed, site, g, L.ptr(t), L.stream_ptr()))
                out[name] = t
        return out

    def ait_forward_train(self, x_props, x_query, token_major_out=False):
        """Transformer.forward keeping the activations the backward needs -> (out, saved buffer).  Engine dtype "tf32" (fp32
        storage, tf32 tensor-core math) or "bf16" (bf16 storage and math; fp32 accumulation, statistics, parameter gradients).
        token_major_out (tf32 only): no NCHW copy -- `out` is the [bp,64,1024] token-major, tf32-rounded result living inside
        the saved buffer (the operand layout of the next stage's GEMMs)."""
        if self.mode not in ("tf32", "bf16"):
            raise RuntimeError("ait_b200: the training path runs in the fp32-storage / tf32 or the bf16 configuration "
                               "(engine dtype 'tf32' | 'bf16'), not %r" % self.mode)
        if token_major_out and self.mode != "tf32":
            raise RuntimeError("ait_b200: the token-major training hand-over exists in the tf32 configuration only")
        lib = L.load()
        ops._need_cuda(x_props, x_query)
        x_props = x_props.contiguous().float()
        x_query = x_query.contiguous().float()
        bp, bs = x_props.shape[0], x_query.shape[0]
        P = bp // bs
        nbytes = lib.aitb_ait_saved_bytes(bs, P)
        saved = torch.empty(int(nbytes) + 1024, dtype=torch.uint8, device=x_props.device)
        sv = saved[(-saved.data_ptr()) % 1024:]
        if token_major_out:
            off = int(lib.aitb_ait_saved_offset(bs, P, 1))
            out = sv[off:off + bp * 64 * 1024 * 4].view(torch.float32).view(bp, 64, 1024)
            out_ptr = L.ptr(None)
        else:
            out = torch.empty((bp, 1024, 8, 8), dtype=torch.float32, device=x_props.device)
            out_ptr = L.ptr(out)
        with torch.cuda.device(x_props.device):
            L.check(lib.aitb_ait_forward_train(C.byref(self.w), L.ptr(x_props), L.ptr(x_query), bs, P, out_ptr,
                                               L.ptr(sv), nbytes, L.stream_ptr()))
        return out, sv

    def ait_backward(self, grad_out, saved, bs, P, token_major_grad=False):
        """-> (grad_props [bp,1024,7,7], grad_query [bs,1024,8,8], [gradient per name of ait_param_names()]).
        token_major_grad: grad_out is [bp,64,1024] token-major and already tf32-rounded (see ait_forward_train)."""
        if token_major_grad and self.mode != "tf32":
            raise RuntimeError("ait_b200: the token-major training hand-over exists in the tf32 configuration only")
        lib = L.load()
        dev = grad_out.device
        grad_out = grad_out.contiguous().float()
        bp = bs * P
        g_props = torch.empty((bp, 1024, 7, 7), dtype=torch.float32, device=dev)
        g_query = torch.empty((bs, 1024, 8, 8), dtype=torch.float32, device=dev)
        shapes = []                      # (struct path, numel) in C-struct order; one flat zeroed buffer behind all
        lin = lambda n, k: [("w", n * k), ("bias", n)]        # noqa: E731
        ln = [("gamma", 512), ("beta", 512)]
        spec = [("enc_emb", lin(512, 1024)), ("dec_emb", lin(512, 1024)), ("dec_trans", lin(1024, 512)),
                ("enc_ln", ln), ("dec_ln", ln)]
        for m in ("enc_slf", "dec_slf", "dec_enc"):
            spec.append((m, [("w_qkv", 1536 * 512), ("w_sk", 512 * 64), ("b_sk", 512), ("w_fc", 512 * 64),
                             ("ln.gamma", 512), ("ln.beta", 512)]))
        for m in ("enc_ffn", "dec_ffn"):
            spec.append((m, [("w1.w", 2048 * 512), ("w1.bias", 2048), ("w2.w", 512 * 2048), ("w2.bias", 512),
                             ("ln.gamma", 512), ("ln.beta", 512)]))
        total = sum(n for _, fs in spec for _, n in fs)
        flat = torch.zeros(total, dtype=torch.float32, device=dev)
        G = L.AITGrads()
        views, off = {}, 0
        for top, fs in spec:
            for path, n in fs:
                obj = getattr(G, top)
                parts = path.split(".")
                for a in parts[:-1]:
                    obj = getattr(obj, a)
                v = flat[off:off + n]
                setattr(obj, parts[-1], v.data_ptr())
                views[top + "." + path] = v
                off += n
        nbytes = lib.aitb_ait_backward_workspace_bytes(bs, P)
        ws = self._workspace(nbytes)
        with torch.cuda.device(dev):
            fn = lib.aitb_ait_backward_tm if token_major_grad else lib.aitb_ait_backward
            L.check(fn(C.byref(self.w), L.ptr(grad_out), bs, P, L.ptr(saved), lib.aitb_ait_saved_bytes(bs, P), C.byref(G),
                       L.ptr(g_props), L.ptr(g_query), L.ptr(ws), nbytes, L.stream_ptr()))
        v = views
        out = [v["enc_emb.w"].view(512, 1024, 1, 1), v["enc_emb.bias"], v["dec_emb.w"].view(512, 1024, 1, 1),
               v["dec_emb.bias"], v["dec_trans.w"].view(1024, 512, 1, 1), v["dec_trans.bias"],
               v["enc_ln.gamma"], v["enc_ln.beta"], v["dec_ln.gamma"], v["dec_ln.beta"]]
        for m in ("enc_slf", "dec_slf", "dec_enc"):
            wqkv = v[m + ".w_qkv"].view(1536, 512)
            out += [wqkv[:512], wqkv[512:1024], wqkv[1024:], v[m + ".w_sk"].view(512, 64), v[m + ".b_sk"],
                    v[m + ".w_fc"].view(512, 64), v[m + ".ln.gamma"], v[m + ".ln.beta"]]
        for m in ("enc_ffn", "dec_ffn"):
            out += [v[m + ".w1.w"].view(2048, 512), v[m + ".w1.bias"], v[m + ".w2.w"].view(512, 2048),
                    v[m + ".w2.bias"], v[m + ".ln.gamma"], v[m + ".ln.beta"]]
        return g_props, g_query, out

    def workspace_bytes(self, B, P):
        return L.load().aitb_head_workspace_bytes(B, P, self.dt)

    # units processed per library call: bounds the activation workspace (~0.87 GB fp32 / 0.44 GB bf16 per
    # unit at P = 300) without changing any result -- units are independent
    MAX_UNITS_PER_CALL = 16

    def head_forward(self, non_img, non_qry, rois, taps=False, out=None):
        B = non_img.shape[0]
        if B > self.MAX_UNITS_PER_CALL and not taps:
            P = rois.shape[1]
            if out is None:
                out = (torch.empty((B, P, 1), dtype=torch.float32, device=non_img.device),
                       torch.empty((B, P, 4), dtype=torch.float32, device=non_img.device))
            for u0 in range(0, B, self.MAX_UNITS_PER_CALL):
                u1 = min(B, u0 + self.MAX_UNITS_PER_CALL)
                r = rois[u0:u1].clone()
                r[..., 0] -= u0                       # roi batch index is relative to the chunk's maps
                self._head_forward_chunk(non_img[u0:u1], non_qry[u0:u1], r, False, (out[0][u0:u1], out[1][u0:u1]))
            return out
        return self._head_forward_chunk(non_img, non_qry, rois, taps, out)

    def _head_forward_chunk(self, non_img, non_qry, rois, taps=False, out=None):
        if not (self.has_ait and self.has_sk and self.has_top and self.has_heads):
            raise RuntimeError("engine built without the full head")
        lib = L.load()
        ops._need_cuda(non_img, non_qry, rois)
        if non_img.dtype != torch.float32 or non_qry.dtype != torch.float32:
            raise RuntimeError("ait_b200: feature maps are passed as float32 NCHW (the compute dtype is an "
                               "engine property)")
        non_img = non_img.contiguous()
        non_qry = non_qry.contiguous()
        B, Cc, H, W = non_img.shape
        if Cc != 1024 or tuple(non_qry.shape) != (B, 1024, 8, 8):
            raise RuntimeError("expected non_img [B,1024,H,W] and non_qry [B,1024,8,8]")
        if rois.dim() != 3 or rois.shape[0] != B or rois.shape[2] != 5:
            raise RuntimeError("expected rois [B, P, 5]")
        P = rois.shape[1]
        rois2 = rois.reshape(B * P, 5).contiguous().float()
        dev = non_img.device
        if out is None:
            cls_prob = torch.empty((B, P, 1), dtype=torch.float32, device=dev)
            bbox = torch.empty((B, P, 4), dtype=torch.float32, device=dev)
        else:
            cls_prob, bbox = out
        tp = None
        tensors = {}
        if taps:
            tp = L.HeadTaps()
            bp = B * P
            pl = self.planes
            tensors = dict(
                pooled=torch.empty((bp, 49, 1024 * pl), dtype=self.dtype, device=dev),
                enc_out=torch.empty((bp, 64, 512 * pl), dtype=self.dtype, device=dev),
                ait_out=torch.empty((bp, 64, 1024 * pl), dtype=self.dtype, device=dev),
                sk_out=torch.empty((bp, 64, 1024 * pl), dtype=self.dtype, device=dev),
                feat=torch.empty((bp, 2048), dtype=torch.float32, device=dev),
                qfeat=torch.empty((B, 2048), dtype=torch.float32, device=dev))
            for k, v in tensors.items():
                setattr(tp, k, v.data_ptr())
        nbytes = lib.aitb_head_workspace_bytes(B, P, self.dt)
        ws = self._workspace(nbytes)
        with torch.cuda.device(dev):
            L.check(lib.aitb_head_forward(C.byref(self.w), L.ptr(non_img), H, W, L.ptr(non_qry), L.ptr(rois2), B, P,
                                          L.ptr(cls_prob), L.ptr(bbox), C.byref(tp) if tp is not None else None,
                                          L.ptr(ws), nbytes, L.stream_ptr()))
        if taps:
            if self.split:
                for k in ("pooled", "enc_out", "ait_out", "sk_out"):
                    tensors[k] = ops.join_planes(tensors[k], f16=self.enc_f16 and k in ("pooled", "enc_out"))
            return cls_prob, bbox, tensors
        return cls_prob, bbox

    # ---- module-level forwards of SKNet / RCNN_top, composed from the exported GEMM building block
    def _sk_branch(self, x_nchw, which):
        m1, b1, m3, b3, mf = self._sk_mats[which]
        G = x_nchw.shape[0]
        sp = self.split
        x = ops.transpose_cs(x_nchw.contiguous().float().reshape(G, 1024, 64), True, out_dtype=self.dtype, split_dst=sp)
        out = torch.empty((G, 64, 1024 * self.planes), dtype=self.dtype, device=x.device)
        # one dual-accumulator GEMM: nine 3x3 taps -> acc0, the 1x1 conv (centre tap) -> acc1
        ops.gemm(x, mf, out, M=G * 64, N=1024, K=128, block_n=128, view="map", map_args=(1024, 8, 8, 1, G),
                 taps=9, group_c=128, flags=L.EPI_BIAS | L.EPI_RELU | L.EPI_SQUARE | L.EPI_DUAL, bias=b3,
                 dual=True, bias2=b1, split=sp)
        return ops.transpose_cs(out, False, out_dtype=torch.float32, split_src=sp).view(G, 1024, 8, 8)

    def sk_forward(self, x_props, x_query):
        if not self.has_sk:
            raise RuntimeError("engine built without SKNet")
        return self._sk_branch(x_props, "props"), self._sk_branch(x_query, "query")

    def top_forward(self, x_nchw):
        """`_head_to_tail`: layer4 + spatial mean, [G,1024,8,8] -> [G,2048]."""
        if not self.has_top:
            raise RuntimeError("engine built without RCNN_top")
        G = x_nchw.shape[0]
        dev = x_nchw.device
        sp, pl = self.split, self.planes
        x = ops.transpose_cs(x_nchw.contiguous().float().reshape(G, 1024, 64), True, out_dtype=self.dtype, split_dst=sp)
        M = G * 16
        c1 = torch.empty((M, 512 * pl), dtype=self.dtype, device=dev)
        c2 = torch.empty((M, 512 * pl), dtype=self.dtype, device=dev)
        ys = [torch.empty((M, 2048 * pl), dtype=self.dtype, device=dev) for _ in range(2)]
        cur = x
        for i, mats in enumerate(self._top_mats):
            out = ys[i & 1]
            w1, b1 = mats["conv1"]
            if i == 0:
                ops.gemm(cur, w1, c1, M=M, N=512, K=1024, block_n=256, view="map", map_args=(1024, 8, 4, 2, G),
                         flags=L.EPI_BIAS | L.EPI_RELU, bias=b1, split=sp)
            else:
                ops.gemm(cur, w1, c1, M=M, N=512, K=2048, block_n=256, flags=L.EPI_BIAS | L.EPI_RELU, bias=b1,
                         split=sp)
            w2, b2 = mats["conv2"]
            ops.gemm(c1, w2, c2, M=M, N=512, K=512, block_n=256, view="map", map_args=(512, 4, 4, 1, G), taps=9,
                     flags=L.EPI_BIAS | L.EPI_RELU, bias=b2, split=sp)
            res = cur
            if i == 0:
                wd, bd = mats["down"]
                ds = torch.empty((M, 2048 * pl), dtype=self.dtype, device=dev)
                ops.gemm(cur, wd, ds, M=M, N=2048, K=1024, block_n=256, view="map", map_args=(1024, 8, 4, 2, G),
                         flags=L.EPI_BIAS, bias=bd, split=sp)
                res = ds
            w3, b3 = mats["conv3"]
            ops.gemm(c2, w3, out, M=M, N=2048, K=512, block_n=256,
                     flags=L.EPI_BIAS | L.EPI_RES | L.EPI_RES_RELU, bias=b3, res=res, ldr=2048, split=sp)
            cur = out
        feat, _, _ = ops.pool_heads(cur.view(G, 16, 2048 * pl), 1, split=sp)
        return feat
