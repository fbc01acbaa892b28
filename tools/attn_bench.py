"""Selective-head attention core at the benchmark shape (2400 pairs, encoder self-attention: 49 keys), hot, CUDA events.
    python tools/attn_bench.py [fp32|bf16 ...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ait_b200 import ops  # noqa: E402

dev = "cuda:0"
G = 2400
for mode in (sys.argv[1:] or ["fp32", "bf16"]):
    split = mode == "fp32"
    conv = (lambda x: ops.split_planes(x)) if split else (lambda x: x.to(torch.bfloat16))
    qkv = conv(torch.randn(G * 64, 1536, device=dev))
    w_sk, b_sk = torch.randn(512, 64, device=dev) * 0.1, torch.zeros(512, device=dev)
    ao = torch.empty(G * 64, 64 * (2 if split else 1), device=dev, dtype=torch.bfloat16)

    def run():
        ops.attn_core(qkv, 1536, 1, qkv.view(-1)[512:], qkv.view(-1)[1024:], 1536, w_sk, b_sk, G, 0, 49, ao, split=split)

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20):
        run()
    e.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(e) / 20 * 1e3
    byt = G * 64 * 1536 * (4 if split else 2) + ao.numel() * 2
    print(json.dumps({"mode": mode, "attn_us": round(us, 1), "GBs": round(byt / us / 1e3, 1)}))
