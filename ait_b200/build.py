"""Build libaitb200.so (sm_100a only) in-tree with nvcc.  Cross-compiles without a GPU.

    python -m ait_b200.build [-v] [--force]

One object per .cu (compiled in parallel), linked into ait_b200/libaitb200.so with the static
CUDA runtime, so the library has no torch / libcudart.so dependency.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libaitb200.so")
SOURCES = ["capi.cu", "gemm.cu", "attn.cu", "nms.cu", "topk.cu", "roi_align.cu", "heads.cu", "bwd.cu", "boxes.cu", "coatt.cu", "attn_tc.cu", "targets.cu", "fc_ln.cu", "heads_train.cu", "train_aux.cu", "dropout.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
]


def _deps_mtime():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "include", "aitb200.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(src, verbose, extra):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    cmd = [NVCC] + FLAGS + extra + ["-c", os.path.join(CSRC, src), "-o", obj]
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    if verbose and (r.stdout or r.stderr):
        print(r.stdout + r.stderr)
    return obj


def build(verbose=False, force=False, ptxas_info=False):
    os.makedirs(OBJ, exist_ok=True)
    hdr_t = _deps_mtime()
    todo, objs = [], []
    for s in SOURCES:
        obj = os.path.join(OBJ, s.replace(".cu", ".o"))
        objs.append(obj)
        src_t = max(os.path.getmtime(os.path.join(CSRC, s)), hdr_t)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < src_t:
            todo.append(s)
    extra = ["-Xptxas", "-v"] if ptxas_info else []
    if todo:
        with ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            list(ex.map(lambda s: _compile(s, verbose or ptxas_info, extra), todo))
    if todo or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                      "-Xcompiler", "-fPIC", "-cudart", "static"]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="--force" in sys.argv, ptxas_info="--ptxas" in sys.argv))
