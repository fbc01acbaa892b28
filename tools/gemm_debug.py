"""Bring-up helper for the tcgen05 GEMM (run on the GPU box): each case runs in its own subprocess so a
trap / launch failure cannot poison the next one.  Prints error statistics and a coarse error map."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [
    # name, dtype, M, N, K, block_n, flags
    ("f32_1chunk_bn64", "f32", 128, 64, 32, 64),
    ("f32_k128_bn64", "f32", 128, 64, 128, 64),
    ("f32_bn128", "f32", 128, 128, 256, 128),
    ("f32_bn256", "f32", 256, 256, 512, 256),
    ("f32_bn256_multi", "f32", 1000, 1536, 512, 256),
    ("f32_bn512", "f32", 256, 512, 256, 512),
    ("bf16_1chunk_bn64", "bf16", 128, 64, 64, 64),
    ("bf16_bn256", "bf16", 256, 256, 512, 256),
    ("bf16_bn512", "bf16", 384, 512, 2048, 512),
    ("f32_big", "f32", 153600, 2048, 512, 256),
]


def run_case(idx):
    import torch
    from ait_b200 import ops
    from ait_b200.packing import round_to_tf32
    name, dt, M, N, K, bn = CASES[idx]
    dtype = torch.float32 if dt == "f32" else torch.bfloat16
    g = torch.Generator().manual_seed(idx)
    big = M > 100000
    dev = "cuda:0"
    if big:
        a = round_to_tf32(torch.randn(M, K, device=dev))
        w = round_to_tf32(torch.randn(N, K, device=dev) / K ** 0.5)
    else:
        a = torch.randn(M, K, generator=g)
        w = torch.randn(N, K, generator=g) / K ** 0.5
        a = round_to_tf32(a) if dt == "f32" else a.to(torch.bfloat16).float()
        w = round_to_tf32(w) if dt == "f32" else w.to(torch.bfloat16).float()
    out = torch.full((M, N), float("nan"), dtype=dtype, device=dev)
    ad, wd = a.to(dev, dtype), w.to(dev, dtype)
    ops.gemm(ad, wd, out, M=M, N=N, K=K, block_n=bn)
    torch.cuda.synchronize()
    if big:
        torch.backends.cuda.matmul.allow_tf32 = False
        ref = ad[:4096].double() @ wd.double().t()
        err = (out[:4096].double() - ref).abs()
        print(name, "max_err(first 4096 rows)", float(err.max()), "nan", int(torch.isnan(out).sum()))
        # timing
        st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            ops.gemm(ad, wd, out, M=M, N=N, K=K, block_n=bn)
        st.record()
        for _ in range(10):
            ops.gemm(ad, wd, out, M=M, N=N, K=K, block_n=bn)
        en.record()
        torch.cuda.synchronize()
        ms = st.elapsed_time(en) / 10
        print(name, "ms", ms, "TFLOP/s", 2.0 * M * N * K / ms / 1e9)
        return
    ref = a.double() @ w.double().t()
    o = out.float().cpu().double()
    err = (o - ref).abs()
    nan = int(torch.isnan(o).sum())
    print(name, "max_err", float(err.nan_to_num(1e9).max()), "ref_max", float(ref.abs().max()), "nan", nan)
    if float(err.nan_to_num(1e9).max()) > 1e-2:
        # coarse map: rows in blocks of 8 (first 64 rows), cols in blocks of 8 (first 64 cols)
        e = err.nan_to_num(9.0)[:64, :64].view(8, 8, 8, 8).amax(dim=(1, 3))
        print("error map (8x8 blocks of the top-left 64x64):")
        for r in e.tolist():
            print(" ".join("%7.3f" % v for v in r))
        print("row0 out:", o[0, :8].tolist())
        print("row0 ref:", ref[0, :8].tolist())


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_case(int(sys.argv[1]))
    else:
        for i in range(len(CASES)):
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), str(i)], capture_output=True,
                                   text=True, timeout=180)
                print(r.stdout.strip())
                if r.returncode != 0:
                    print(CASES[i][0], "FAILED rc", r.returncode, r.stderr.strip()[-600:])
            except subprocess.TimeoutExpired:
                print(CASES[i][0], "TIMEOUT")
