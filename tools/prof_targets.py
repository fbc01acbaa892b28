"""Launch the hot kernels in isolation on benchmark-shaped data (for `ncu --set full`):
FFN w_1 GEMM (2-CTA), FFN w_2 GEMM (cluster LayerNorm), attention fc GEMM (K = 64, LayerNorm) and its streaming replacement (fc_ln), encoder
self-attention core, ROIAlign (token-major), proposal top-n + NMS.

    python tools/prof_targets.py [reps] [fp32|tf32|bf16]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ait_b200 import _lib as L, ops, synth  # noqa: E402
from ait_b200.proposal import propose_rois  # noqa: E402

dev = "cuda:0"
B, P = 8, 300
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
mode = sys.argv[2] if len(sys.argv) > 2 else "fp32"
split = mode == "fp32"
dt = torch.bfloat16 if mode == "bf16" else torch.float32


def act(x):
    """fp32 tensor -> activation / weight buffer of the mode."""
    return ops.split_planes(x) if split else x.to(dt)


def buf(rows, cols):
    return torch.empty(rows, cols * (2 if split else 1), device=dev, dtype=torch.bfloat16 if split else dt)


# ROIAlign inputs
maps = torch.stack([synth.c4_map(u) for u in range(B)]).to(dev)
rpn = [synth.rpn_outputs(u) for u in range(B)]
boxes, scores = torch.stack([r[0] for r in rpn]).to(dev), torch.stack([r[1] for r in rpn]).to(dev)
rois, _ = propose_rois(boxes, scores)
nhwc = ops.transpose_cs(maps.reshape(B, 1024, -1), True, out_dtype=dt).view(B, 38, 63, 1024)
# GEMM inputs
M = B * P * 64
x512 = act(torch.randn(M, 512, device=dev))
x64 = act(torch.randn(M, 64, device=dev))
w1 = act(torch.randn(2048, 512, device=dev) / 512 ** 0.5)
w2 = act(torch.randn(512, 2048, device=dev) / 2048 ** 0.5)
wfc = act(torch.randn(512, 64, device=dev) / 8)
hid = buf(M, 2048)
o512 = buf(M, 512)
b2048 = torch.zeros(2048, device=dev)
b512 = torch.zeros(512, device=dev)
gamma = torch.ones(512, device=dev)
# attention inputs
G = B * P
qkv = act(torch.randn(G * 64, 1536, device=dev))
w_sk = torch.randn(512, 64, device=dev) * 0.1
b_sk = torch.zeros(512, device=dev)
ao = buf(G * 64, 64)
cb = 1 if split else 1  # column offsets below are in storage elements of the hi plane
sk_in = act(torch.randn(G * 64, 1024, device=dev)).view(G, 64, -1)
sk_w = act(torch.randn(1024, 1280, device=dev) / 1280 ** 0.5)
sk_out = buf(G * 64, 1024).view(G, 64, -1)
b1024 = torch.zeros(1024, device=dev)
if split:
    x512_16 = ops.split_planes(torch.randn(M, 512, device=dev), f16=True)
    wqkv_16 = ops.split_planes(torch.randn(1536, 512, device=dev) / 512 ** 0.5, f16=True)
    w2_16 = ops.split_planes(torch.randn(512, 2048, device=dev) / 2048 ** 0.5, f16=True)
    qkv_out = buf(M, 1536)
# the nms(dets, scores, thr) drop-in on one image's top-6000 boxes (bitmask + on-device scan)
from ait_b200.roi_layers import nms as nms_dropin  # noqa: E402
o0 = torch.argsort(scores[0], descending=True)[:6000]
d0, s0 = boxes[0][o0].contiguous(), scores[0][o0].contiguous()
torch.cuda.synchronize()
for _ in range(reps):
    nms_dropin(d0, s0, 0.7)
    ops.gemm(x512, w1, hid, M=M, N=2048, K=512, block_n=256, flags=L.EPI_BIAS | L.EPI_RELU, bias=b2048, split=split)
    ops.gemm(hid, w2, o512, M=M, N=512, K=2048, block_n=512, flags=L.EPI_BIAS | L.EPI_RES | L.EPI_LN, bias=b512,
             res=x512, ldr=512, gamma=gamma, beta=b512, split=split)
    ops.gemm(x64, wfc, o512, M=M, N=512, K=64, block_n=512, flags=L.EPI_RES | L.EPI_LN, res=x512, ldr=512,
             res_div=64, gamma=gamma, beta=b512, split=split)
    if mode != "tf32":   # the streaming fc + residual + LayerNorm kernel that replaced the K = 64 GEMM above on the hot path
        ops.fc_ln(x64, wfc, x512, gamma, b512, o512, M=M, split=split, res_div=64, res_rep=1)
    ops.attn_core(qkv, 1536, 1, qkv.view(-1)[512:], qkv.view(-1)[1024:], 1536, w_sk, b_sk, G, 0, 49, ao, split=split)
    if split:
        lib = L.load()
        pooled = torch.empty(G, 49, 2048, device=dev, dtype=torch.bfloat16)
        L.check(lib.aitb_roi_align_forward(L.ptr(nhwc), L.ptr(rois.view(-1, 5)), B, 1024, 38, 63, G, 1 / 16.0, 7, 7, 0,
                                           L.AITB_F32S, 1, L.ptr(pooled), L.stream_ptr()))
    else:
        ops.roi_align_forward(nhwc, rois.view(-1, 5), 1 / 16.0, 7, 7, 0, token_major=True)
    propose_rois(boxes, scores)
    # the grouped SKBlock convolutions as one dual-accumulator GEMM (gemm_tcgen05_kernel<.., 128, 0, SPLIT>: 9 shifted 3x3 taps + the
    # 1x1 centre tap, relu^2 sum epilogue) on the proposal maps [G, 8, 8, 1024]
    ops.gemm(sk_in, sk_w, sk_out, M=G * 64, N=1024, K=128, block_n=128, view="map", map_args=(1024, 8, 8, 1, G), taps=9,
             group_c=128, flags=L.EPI_BIAS | L.EPI_RELU | L.EPI_SQUARE | L.EPI_DUAL, bias=b1024, dual=True, bias2=b1024, split=split)
    if split:   # precision plan: the one-pass variants (fp16 hi planes) of the encoder QKV projection and the FFN w_2 GEMM
        ops.gemm(x512_16, wqkv_16, qkv_out, M=M, N=1536, K=512, block_n=256, split=True, passes=1, in_f16=True)
        ops.gemm(hid, w2_16, o512, M=M, N=512, K=2048, block_n=512, flags=L.EPI_BIAS | L.EPI_RES | L.EPI_LN, bias=b512,
                 res=x512_16, ldr=512, gamma=gamma, beta=b512, split=True, passes=1, in_f16=True, out_f16=True, res_f16=True)
torch.cuda.synchronize()
print("done")
