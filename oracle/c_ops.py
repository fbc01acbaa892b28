"""TEST INFRASTRUCTURE ONLY.  ctypes front-end of oracle/oracle_ops.c (built on demand with gcc into
oracle/_build/, which travels to the GPU box like any other built .so)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "oracle_ops.c")
OUT_DIR = os.path.join(HERE, "_build")
SO = os.path.join(OUT_DIR, "liboracle_ops.so")
_lib = None


def build():
    os.makedirs(OUT_DIR, exist_ok=True)
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", SRC, "-o", SO, "-lm"])
    return SO


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_nms_sorted.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def nms_sorted(boxes_sorted, thr, ge=False, max_keep=0):
    """boxes_sorted [n,4] float32 (descending score order) -> kept positions (int32, ascending)."""
    b = np.ascontiguousarray(boxes_sorted, dtype=np.float32)
    n = b.shape[0]
    keep = np.empty(max(n, 1), dtype=np.int32)
    k = lib().oracle_nms_sorted(_p(b), C.c_int(n), C.c_float(thr), C.c_int(1 if ge else 0), C.c_int(max_keep),
                                _p(keep))
    return keep[:k].copy()


def nms(dets, scores, thr, ge=False):
    """The reference `nms(dets, scores, thr)` contract: kept ORIGINAL indices, ascending (nms.cu:127-130).
    Order = descending score, ties by lower index (stable)."""
    dets = np.ascontiguousarray(dets, dtype=np.float32)
    scores = np.asarray(scores, dtype=np.float32)
    if dets.shape[0] == 0:
        return np.empty(0, dtype=np.int64)
    order = np.argsort(-scores.astype(np.float64), kind="stable")
    kept = nms_sorted(dets[order], thr, ge=ge)
    return np.sort(order[kept]).astype(np.int64)


def roi_align_forward(feat, rois, scale, ph, pw, sampling_ratio):
    feat = np.ascontiguousarray(feat, dtype=np.float32)
    rois = np.ascontiguousarray(rois, dtype=np.float32)
    B, Cc, H, W = feat.shape
    K = rois.shape[0]
    out = np.empty((K, Cc, ph, pw), dtype=np.float32)
    lib().oracle_roi_align_forward(_p(feat), _p(rois), C.c_int(Cc), C.c_int(H), C.c_int(W), C.c_int(K),
                                   C.c_float(scale), C.c_int(ph), C.c_int(pw), C.c_int(sampling_ratio), _p(out))
    return out


def roi_align_backward(grad, rois, scale, ph, pw, B, Cc, H, W, sampling_ratio):
    grad = np.ascontiguousarray(grad, dtype=np.float32)
    rois = np.ascontiguousarray(rois, dtype=np.float32)
    K = rois.shape[0]
    g = np.zeros((B, Cc, H, W), dtype=np.float64)
    lib().oracle_roi_align_backward(_p(grad), _p(rois), C.c_int(Cc), C.c_int(H), C.c_int(W), C.c_int(K),
                                    C.c_float(scale), C.c_int(ph), C.c_int(pw), C.c_int(sampling_ratio), _p(g))
    return g
