"""Time every distinct GEMM launch of one benchmark step (8 units x 300 proposals) in isolation:
    python tools/step_gemms.py [fp32|tf32|bf16] [name-filter]
Reports ms, algorithmic TFLOP/s, fraction of the configuration's tensor peak, and the minimum HBM GB/s."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ait_b200 import _lib as L, ops  # noqa: E402

dev = "cuda:0"
mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
filt = sys.argv[2] if len(sys.argv) > 2 else ""
split = mode == "fp32"
dt = torch.float32 if mode == "tf32" else torch.bfloat16
pl = 2 if split else 1
eb = 2 if mode == "bf16" else 4
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
peak = peaks["bf16_tflops"] / {"bf16": 1, "tf32": 2, "fp32": 3}[mode]
BP = 2400
R = BP * 64


def act(rows, cols):
    x = torch.randn(rows, cols, device=dev)
    return ops.split_planes(x) if split else x.to(dt)


def empty(rows, cols):
    return torch.empty(rows, cols * pl, device=dev, dtype=dt)


def vec(n):
    return torch.randn(n, device=dev)


def run(name, M, N, K, taps, fn, bytes_min):
    if filt and filt not in name:
        return
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st.record()
    for _ in range(5):
        fn()
    en.record()
    torch.cuda.synchronize()
    ms = st.elapsed_time(en) / 5
    tf = 2.0 * M * N * K * taps / ms / 1e9
    print("%-4s %-22s M=%6d N=%4d K=%4dx%d  %7.3f ms  %6.0f TF/s  %4.0f%% of peak   min-HBM %5.0f GB/s"
          % (mode, name, M, N, K, taps, ms, tf, 100 * tf / peak, bytes_min / ms / 1e6))


def plain(name, M, N, K, flags, res=False, ln=False, rows=None):
    a, w, o = act(M, K), act(N, K), empty(M if rows is None else rows, N)
    bias = vec(N) if flags & L.EPI_BIAS else None
    r = act(M, N) if res else None
    g, b = (vec(N), vec(N)) if ln else (None, None)
    pos = torch.randn(64, N, device=dev) if flags & L.EPI_POS else None
    kw = {}
    if rows is not None:
        kw = dict(rows_in=49, rows_out=64)
    fn = lambda: ops.gemm(a, w, o, M=M, N=N, K=K, block_n=512 if ln else 256, flags=flags, bias=bias, res=r, ldr=N,  # noqa: E731
                          gamma=g, beta=b, pos=pos, pos_rows=64, split=split, **kw)
    run(name, M, N, K, 1, fn, (M * K + M * N * (2 if res else 1)) * eb)


plain("enc_emb+pos+LN", BP * 49, 512, 1024, L.EPI_BIAS | L.EPI_POS | L.EPI_LN, ln=True, rows=R)
plain("qkv", R, 1536, 512, 0)
plain("attn_fc+res+LN", R, 512, 64, L.EPI_RES | L.EPI_LN, res=True, ln=True)
plain("ffn_w1+relu", R, 2048, 512, L.EPI_BIAS | L.EPI_RELU)
plain("ffn_w2+res+LN", R, 512, 2048, L.EPI_BIAS | L.EPI_RES | L.EPI_LN, res=True, ln=True)
plain("kv", R, 1024, 512, 0)
plain("dec_trans+bias", R, 1024, 512, L.EPI_BIAS)
M16 = BP * 16
plain("l4_conv3+res+relu", M16, 2048, 512, L.EPI_BIAS | L.EPI_RES | L.EPI_RES_RELU, res=True)
plain("l4_conv3 (no res)", M16, 2048, 512, L.EPI_BIAS)
plain("l4_conv1 K2048", M16, 512, 2048, L.EPI_BIAS | L.EPI_RELU)

# map views
x8 = act(BP * 64, 1024)
w_c1 = act(512, 1024)
o_c1 = empty(M16, 512)
b512 = vec(512)
run("l4_conv1 stride2", M16, 512, 1024, 1,
    lambda: ops.gemm(x8, w_c1, o_c1, M=M16, N=512, K=1024, block_n=256, view="map", map_args=(1024, 8, 4, 2, BP),
                     flags=L.EPI_BIAS | L.EPI_RELU, bias=b512, split=split), (M16 * 1024 + M16 * 512) * eb)
w_dn = act(2048, 1024)
o_dn = empty(M16, 2048)
b2048 = vec(2048)
run("l4_down stride2", M16, 2048, 1024, 1,
    lambda: ops.gemm(x8, w_dn, o_dn, M=M16, N=2048, K=1024, block_n=256, view="map", map_args=(1024, 8, 4, 2, BP),
                     flags=L.EPI_BIAS, bias=b2048, split=split), (M16 * 1024 + M16 * 2048) * eb)
x4 = act(M16, 512)
w_c2 = act(512, 9 * 512)
o_c2 = empty(M16, 512)
run("l4_conv2 3x3", M16, 512, 512, 9,
    lambda: ops.gemm(x4, w_c2, o_c2, M=M16, N=512, K=512, block_n=256, view="map", map_args=(512, 4, 4, 1, BP), taps=9,
                     flags=L.EPI_BIAS | L.EPI_RELU, bias=b512, split=split), (M16 * 512 * 2) * eb)
w_sk = act(1024, 10 * 128)
o_sk = empty(R, 1024)
b1024, b1024b = vec(1024), vec(1024)
run("sk dual grouped", R, 1024, 128, 10,
    lambda: ops.gemm(x8, w_sk, o_sk, M=R, N=1024, K=128, block_n=128, view="map", map_args=(1024, 8, 8, 1, BP), taps=9,
                     group_c=128, flags=L.EPI_BIAS | L.EPI_RELU | L.EPI_SQUARE | L.EPI_DUAL, bias=b1024, dual=True,
                     bias2=b1024b, split=split), (R * 1024 * 2) * eb)
# epilogue-only probes: one K chunk, so the time is TMEM drain + epilogue math + global stores
plain("probe K64 N2048 bias+relu", R, 2048, 64, L.EPI_BIAS | L.EPI_RELU)
plain("probe K64 N2048 plain", R, 2048, 64, 0)
plain("probe K64 N2048 +res", R, 2048, 64, L.EPI_RES, res=True)
plain("probe K128 N2048 plain", R, 2048, 128, 0)
plain("probe K256 N2048 plain", R, 2048, 256, 0)
plain("probe K1024 N2048 plain", R, 2048, 1024, 0)
