"""CPU simulation of a per-stage tensor-core precision plan for the fp32 configuration.

The fp32-class configuration executes every product as three bf16 MMA passes (hi*hi + hi*lo + lo*hi).  This tool
emulates, on the CPU oracle, what each GEMM stage of the head would contribute to the final cls_prob / feature error if
its operands were rounded more coarsely (= fewer passes), everything else exact:

    p1b   one pass, both operands bf16                     (A 8 bit,  W 8 bit)
    p1h   one pass, both operands 11-bit (fp16 / tf32)     (A 11 bit, W 11 bit)
    p2aw8 two passes  hi*hi + lo*hi, bf16 planes           (A 16 bit, W 8 bit)
    p2a8w two passes  hi*hi + hi*lo, bf16 planes           (A 8 bit,  W 16 bit)
    p2h   two passes  hi*hi + lo*hi, fp16 planes           (A 22 bit, W 11 bit)
    p3    three passes, bf16 planes (today)                (A 16 bit, W 16 bit)

Accumulation is exact (fp32/fp64 on the CPU); the tensor core's fp32 accumulation error is common to all variants.
The reference is the fp64 oracle.  Stages are the head's GEMMs in call order (oracle/head_oracle.py).

    python tools/precision_plan.py [--props 16] [--modes p1h,p2aw8]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.nn.functional as TF  # noqa: E402

from conftest import golden_head, head_inputs  # noqa: E402
from oracle import head_oracle  # noqa: E402


def rnd_bits(x, bits):
    """round-to-nearest to `bits` significand bits (incl. the hidden one), no range limit."""
    if bits >= 24 and x.dtype == torch.float32:
        return x
    xd = x.double()
    m, e = torch.frexp(xd)
    s = 2.0 ** bits
    return torch.ldexp(torch.round(m * s) / s, e).to(x.dtype)


MODES = {  # (A bits, W bits)
    "p1b": (8, 8), "p1h": (11, 11), "p2aw8": (16, 8), "p2a8w": (8, 16), "p2h": (22, 11), "p3": (16, 16), "exact": (53, 53),
}

# stage names in the order head_oracle.head_forward issues its linear / conv2d / matmul calls
STAGES = [
    "enc_emb", "dec_emb",
    "enc_q", "enc_k", "enc_v", "enc_qk", "enc_pv", "enc_sk", "enc_fc", "enc_w1", "enc_w2",
    "dslf_q", "dslf_k", "dslf_v", "dslf_qk", "dslf_pv", "dslf_sk", "dslf_fc",
    "x_q", "x_k", "x_v", "x_qk", "x_pv", "x_sk", "x_fc", "dec_w1", "dec_w2", "dec_trans",
    "skp_1x1", "skp_3x3", "skq_1x1", "skq_3x3",
] + ["l4p_%s" % n for n in ("b0c1", "b0c2", "b0c3", "b0ds", "b1c1", "b1c2", "b1c3", "b2c1", "b2c2", "b2c3")] \
  + ["l4q_%s" % n for n in ("b0c1", "b0c2", "b0c3", "b0ds", "b1c1", "b1c2", "b1c3", "b2c1", "b2c2", "b2c3")] \
  + ["bbox", "cls1", "cls2"]

GROUPS = {  # engine GEMM launches -> oracle stages
    "enc_emb": ["enc_emb"], "enc_qkv": ["enc_q", "enc_k", "enc_v"], "enc_attn": ["enc_qk", "enc_pv"], "enc_fc": ["enc_fc"],
    "enc_w1": ["enc_w1"], "enc_w2": ["enc_w2"], "x_kv": ["x_k", "x_v"], "x_attn": ["x_qk", "x_pv"], "x_fc": ["x_fc"],
    "dec_w1": ["dec_w1"], "dec_w2": ["dec_w2"], "dec_trans": ["dec_trans"], "sk_props": ["skp_1x1", "skp_3x3"],
    "l4_b0c1": ["l4p_b0c1"], "l4_b0c2": ["l4p_b0c2"], "l4_b0c3": ["l4p_b0c3"], "l4_b0ds": ["l4p_b0ds"],
    "l4_b1c1": ["l4p_b1c1"], "l4_b1c2": ["l4p_b1c2"], "l4_b1c3": ["l4p_b1c3"],
    "l4_b2c1": ["l4p_b2c1"], "l4_b2c2": ["l4p_b2c2"], "l4_b2c3": ["l4p_b2c3"],
}
MFLOP = {  # per pair, for the cost column
    "enc_emb": 51.4, "enc_qkv": 100.7, "enc_attn": 8.4, "enc_fc": 4.2, "enc_w1": 102.8, "enc_w2": 102.8, "x_kv": 51.4,
    "x_attn": 8.4, "x_fc": 4.2, "dec_w1": 134.2, "dec_w2": 134.2, "dec_trans": 67.1, "sk_props": 167.8,
    "l4_b0c1": 16.8, "l4_b0c2": 75.5, "l4_b0c3": 33.6, "l4_b0ds": 67.1, "l4_b1c1": 33.6, "l4_b1c2": 75.5, "l4_b1c3": 33.6,
    "l4_b2c1": 33.6, "l4_b2c2": 75.5, "l4_b2c3": 33.6,
}


class Hook:
    """Stands in for torch.nn.functional / torch inside head_oracle: rounds the operands of selected stages."""

    def __init__(self):
        self.plan = {}
        self.i = 0
        self.fold_bn = None

    def reset(self, plan):
        self.plan, self.i = plan, 0

    def _ops(self, a, w):
        name = STAGES[self.i]
        self.i += 1
        mode = self.plan.get(name)
        if mode is None:
            return a, w
        ab, wb = MODES[mode]
        return rnd_bits(a, ab), rnd_bits(w, wb)

    def linear(self, x, w, b=None):
        x, w = self._ops(x, w)
        return TF.linear(x, w, b)

    def conv2d(self, x, w, b=None, **kw):
        x, w = self._ops(x, w)
        return TF.conv2d(x, w, b, **kw)

    def matmul(self, a, b):
        a, b = self._ops(a, b)
        return torch.matmul(a, b)

    def __getattr__(self, k):
        return getattr(TF, k)


class TorchProxy:
    def __init__(self, hook):
        self.hook = hook

    def matmul(self, a, b):
        return self.hook.matmul(a, b)

    def __getattr__(self, k):
        return getattr(torch, k)


def fold_bn_state(sd):
    """the engine folds the frozen BatchNorm into the conv weights BEFORE splitting them: do the same, so that rounding
    the 'weight' of a layer4 stage rounds the folded weight (BatchNorm becomes identity + bias)."""
    out = dict(sd)
    for k in list(sd):
        if k.startswith("RCNN_top.") and k.endswith("running_var"):
            p = k[: -len("running_var")]                       # RCNN_top.0.<blk>.bnX.  | ...downsample.1.
            conv = p.replace("bn1.", "conv1.").replace("bn2.", "conv2.").replace("bn3.", "conv3.").replace(
                "downsample.1.", "downsample.0.")
            scale = sd[p + "weight"].double() / torch.sqrt(sd[p + "running_var"].double() + 1e-5)
            out[conv + "weight"] = (sd[conv + "weight"].double() * scale.view(-1, 1, 1, 1))
            out[p + "bias"] = sd[p + "bias"].double() - sd[p + "running_mean"].double() * scale
            out[p + "weight"] = torch.ones_like(scale)
            out[p + "running_mean"] = torch.zeros_like(scale)
            out[p + "running_var"] = torch.ones_like(scale) - 1e-5
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--props", type=int, default=16)
    ap.add_argument("--modes", default="p1b,p1h,p2aw8,p2a8w,p2h,p3")
    ap.add_argument("--plan", default="", help="evaluate one full plan: group=mode,... (others p3)")
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    B, P = 2, args.props
    head, g = golden_head()
    sd = fold_bn_state({k: v.clone() for k, v in head.state_dict().items()})
    non_img, non_qry, rois = head_inputs(B, P)
    hook = Hook()
    head_oracle.F = hook
    head_oracle.torch = TorchProxy(hook)

    def run(plan):
        hook.reset(plan)
        with torch.no_grad():
            r = head_oracle.head_forward(sd, non_img, non_qry, rois, dtype=torch.float64)
        assert hook.i == len(STAGES), (hook.i, len(STAGES))
        return r

    ref = run({})
    # stress calibration of the score layer on THESE features (tests/golden/make_golden.py recipe)
    stack = torch.cat([ref["feat"].view(B, P, -1), ref["qfeat"].unsqueeze(1).repeat(1, P, 1)], 2).view(-1, 4096)
    Xc = stack - stack.mean(0, keepdim=True)
    _, S, Vt = torch.linalg.svd(Xc, full_matrices=False)
    A = torch.zeros(8, 4096, dtype=torch.float64)
    for j in range(3):
        A[j] = Vt[j] / (S[j] / 8 ** 0.5)
    b1 = -(stack @ A.t()).mean(0)
    wv = torch.tensor([0.9, -0.7, 0.5, 0, 0, 0, 0, 0], dtype=torch.float64)
    W2 = torch.stack([-wv, wv])

    def probs(r):
        st = torch.cat([r["feat"].view(B, P, -1), r["qfeat"].unsqueeze(1).repeat(1, P, 1)], 2).view(-1, 4096)
        return torch.softmax((st @ A.t() + b1) @ W2.t(), 1)[:, 1]

    p_ref = probs(ref)
    print("calibrated cls_prob span: %.3f .. %.3f" % (float(p_ref.min()), float(p_ref.max())))

    def report(tag, r, cost=None):
        e_p = float((probs(r) - p_ref).abs().max())
        e_f = float((r["feat"] - ref["feat"]).abs().max() / ref["feat"].abs().max())
        e_a = float((r["ait_out"] - ref["ait_out"]).abs().max() / ref["ait_out"].abs().max())
        e_b = float((r["bbox_pred"] - ref["bbox_pred"]).abs().max() / ref["bbox_pred"].abs().max())
        print("%-28s cls_prob %.2e  feat %.2e  ait %.2e  bbox %.2e%s" % (tag, e_p, e_f, e_a, e_b,
                                                                         "" if cost is None else "   passes-weighted MFLOP %.0f" % cost))
        return e_p

    all_groups = list(GROUPS)
    passes = {"p1b": 1, "p1h": 1, "p2aw8": 2, "p2a8w": 2, "p2h": 2, "p3": 3}

    def full(plan_groups):
        plan = {}
        for grp, mode in plan_groups.items():
            for s in GROUPS[grp]:
                plan[s] = mode
        # the query side (tiny) and the heads stay p3 / fp32
        for s in STAGES:
            if s not in plan and s not in ("bbox", "cls1", "cls2", "enc_sk", "dslf_sk", "x_sk"):
                plan[s] = "p3"
        cost = sum(MFLOP[g_] * passes[m] for g_, m in plan_groups.items())
        return plan, cost

    base_plan, base_cost = full({g_: "p3" for g_ in all_groups})
    report("all p3 (today)", run(base_plan), base_cost)
    if args.plan:
        pg = {g_: "p3" for g_ in all_groups}
        for item in args.plan.split(","):
            k, v = item.split("=")
            for g_ in all_groups:
                if g_ == k or (k.endswith("*") and g_.startswith(k[:-1])):
                    pg[g_] = v
        plan, cost = full(pg)
        report("plan " + args.plan, run(plan), cost)
        return
    for mode in args.modes.split(","):
        plan, cost = full({g_: mode for g_ in all_groups})
        report("ALL %s" % mode, run(plan), cost)
    for mode in args.modes.split(","):
        if mode == "p3":
            continue
        print("---- one group at %s, the rest p3" % mode)
        for grp in all_groups:
            pg = {g_: "p3" for g_ in all_groups}
            pg[grp] = mode
            plan, cost = full(pg)
            report("%s=%s" % (grp, mode), run(plan), cost)


if __name__ == "__main__":
    main()
