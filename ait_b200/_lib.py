"""ctypes binding of libaitb200.so (include/aitb200.h).  No fallback: a missing library or a
non-sm_100 device raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libaitb200.so")

AITB_F32, AITB_BF16, AITB_F32S = 0, 1, 2

# Compute configurations of the engine (include/aitb200.h):
#   "fp32"  AITB_F32S  split bf16 hi/lo planes, three bf16 tensor-core passes per product (fp32-class results:
#                      the configuration that meets the reference's fp32 scores to 1e-3)
#   "tf32"  AITB_F32   fp32 storage, tf32 tensor-core math (faster, 10-bit operand mantissa)
#   "bf16"  AITB_BF16  bf16 storage and math
MODES = {"fp32": AITB_F32S, "tf32": AITB_F32, "bf16": AITB_BF16}

PLAN_ENC_ONEPASS = 1

EPI_BIAS, EPI_RELU, EPI_SQUARE, EPI_RES, EPI_POS, EPI_LN, EPI_ACCUM, EPI_RES_RELU, EPI_DUAL, EPI_RELU_MASK, EPI_RES_ROW_M, EPI_HI_ONLY = (
    1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048)


class View4(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("dims", C.c_uint64 * 4), ("strides", C.c_uint64 * 3),
                ("box", C.c_uint32 * 4)]


class GemmDesc(C.Structure):
    _fields_ = [
        ("dtype", C.c_int), ("M", C.c_int), ("N", C.c_int), ("k_per_tap", C.c_int), ("taps", C.c_int),
        ("a", View4), ("a_m_dim", C.c_int), ("a_m_step", C.c_int), ("a_group_c", C.c_int),
        ("tap_dx", C.c_int8 * 9), ("tap_dy", C.c_int8 * 9),
        ("w", C.c_void_p), ("block_n", C.c_int),
        ("flags", C.c_int), ("out", C.c_void_p), ("ldo", C.c_int),
        ("rows_in", C.c_int), ("rows_out", C.c_int),
        ("bias", C.c_void_p), ("res", C.c_void_p), ("ldr", C.c_int),
        ("res_div", C.c_int), ("res_rep", C.c_int),
        ("pos", C.c_void_p), ("pos_rows", C.c_int),
        ("gamma", C.c_void_p), ("beta", C.c_void_p), ("eps", C.c_float), ("round_tf32", C.c_int),
        ("dual", C.c_int), ("bias2", C.c_void_p), ("a_lo_off", C.c_int), ("ln_rstd", C.c_void_p), ("map_w", C.c_int), ("map_h", C.c_int), ("out_scale", C.c_float),
        ("passes", C.c_int), ("in_f16", C.c_int), ("out_f16", C.c_int), ("res_f16", C.c_int),
    ]


class Linear(C.Structure):
    _fields_ = [("w", C.c_void_p), ("bias", C.c_void_p)]


class LNorm(C.Structure):
    _fields_ = [("gamma", C.c_void_p), ("beta", C.c_void_p)]


class MHA(C.Structure):
    _fields_ = [("w_qkv", C.c_void_p), ("w_sk", C.c_void_p), ("b_sk", C.c_void_p), ("w_fc", C.c_void_p),
                ("ln", LNorm)]


class FFN(C.Structure):
    _fields_ = [("w1", Linear), ("w2", Linear), ("ln", LNorm)]


class Bottleneck(C.Structure):
    _fields_ = [("conv1", Linear), ("conv2", Linear), ("conv3", Linear), ("down", Linear)]


class SKBlock(C.Structure):
    _fields_ = [("conv1x1", Linear), ("conv3x3", Linear), ("w_fused", C.c_void_p)]


class HeadWeights(C.Structure):
    _fields_ = [
        ("dtype", C.c_int), ("round_tf32", C.c_int), ("plan", C.c_int),
        ("enc_emb", Linear), ("dec_emb", Linear), ("dec_trans", Linear),
        ("enc_pos", C.c_void_p), ("dec_pos", C.c_void_p),
        ("enc_ln", LNorm), ("dec_ln", LNorm),
        ("enc_slf", MHA), ("dec_slf", MHA), ("dec_enc", MHA),
        ("enc_ffn", FFN), ("dec_ffn", FFN),
        ("sk_props", SKBlock), ("sk_query", SKBlock),
        ("top", Bottleneck * 3),
        ("w_bbox", C.c_void_p), ("b_bbox", C.c_void_p),
        ("w_cls1", C.c_void_p), ("b_cls1", C.c_void_p),
        ("w_cls2", C.c_void_p), ("b_cls2", C.c_void_p),
        ("p_drop", C.c_float), ("p_attn", C.c_float), ("drop_seed", C.c_ulonglong),
    ]


class LinearG(C.Structure):
    _fields_ = [("w", C.c_void_p), ("bias", C.c_void_p)]


class MHAG(C.Structure):
    _fields_ = [("w_qkv", C.c_void_p), ("w_sk", C.c_void_p), ("b_sk", C.c_void_p), ("w_fc", C.c_void_p),
                ("ln", LNorm)]


class FFNG(C.Structure):
    _fields_ = [("w1", LinearG), ("w2", LinearG), ("ln", LNorm)]


class AITGrads(C.Structure):
    """aitb_ait_grads: fp32 gradient buffers of the AIT parameters (include/aitb200.h)."""
    _fields_ = [("enc_emb", LinearG), ("dec_emb", LinearG), ("dec_trans", LinearG),
                ("enc_ln", LNorm), ("dec_ln", LNorm),
                ("enc_slf", MHAG), ("dec_slf", MHAG), ("dec_enc", MHAG),
                ("enc_ffn", FFNG), ("dec_ffn", FFNG)]


class HeadTaps(C.Structure):
    _fields_ = [("pooled", C.c_void_p), ("enc_out", C.c_void_p), ("ait_out", C.c_void_p),
                ("sk_out", C.c_void_p), ("feat", C.c_void_p), ("qfeat", C.c_void_p)]


# every symbol include/aitb200.h declares: (restype, argtypes)
_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
SIGNATURES = {
    "aitb_last_error": (C.c_char_p, []),
    "aitb_version": (_i, []),
    "aitb_check_device": (_i, []),
    "aitb_launch_count": (C.c_longlong, [_i]),
    "aitb_nms_workspace_bytes": (_sz, [_i, _i, _i]),
    "aitb_nms_batched": (_i, [_vp, _vp, _i, _i, _i, _f, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "aitb_topk_workspace_bytes": (_sz, [_i, _i, _i]),
    "aitb_topk_desc": (_i, [_vp, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "aitb_roi_align_forward": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _f, _i, _i, _i, _i, _i, _vp, _vp]),
    "aitb_roi_align_backward": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _f, _i, _i, _i, _vp, _vp]),
    "aitb_transpose_cs": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _i, _vp]),
    "aitb_transpose_cs_round": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "aitb_gemm": (_i, [C.POINTER(GemmDesc), _vp]),
    "aitb_attn_core": (_i, [_vp, _i, _i, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "aitb_pool_heads": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "aitb_head_workspace_bytes": (_sz, [_i, _i, _i]),
    "aitb_head_forward": (_i, [C.POINTER(HeadWeights), _vp, _i, _i, _vp, _vp, _i, _i, _vp, _vp,
                               C.POINTER(HeadTaps), _vp, _sz, _vp]),
    "aitb_ait_workspace_bytes": (_sz, [_i, _i, _i]),
    "aitb_ait_forward": (_i, [C.POINTER(HeadWeights), _vp, _vp, _i, _i, _vp, _vp, _sz, _vp]),
    "aitb_coattention_workspace_bytes": (_sz, [_i, _i, _i]),
    "aitb_coattention_forward": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "aitb_rpn_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "aitb_rpn_forward": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "aitb_rpn_decode": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp, _vp, _vp]),
    "aitb_box_decode": (_i, [_vp, _i, _i, _vp, _vp, _vp, _i, _i, C.POINTER(C.c_float), C.POINTER(C.c_float), _f, _i,
                             _vp, _vp, _vp, _vp]),
    "aitb_det_assemble": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "aitb_wgrad": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _vp, _i, _vp]),
    "aitb_group_norm_forward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, C.c_float, _vp, _vp, _vp]),
    "aitb_group_norm_backward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, C.c_float, _i, _vp, _vp, _vp, _vp, _vp]),
    "aitb_wgrad_bf16": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _vp, _i, _vp]),
    "aitb_wgrad_conv": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _i, _i, _vp, _i, _vp]),
    "aitb_ln_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "aitb_colsum": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "aitb_bsum": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "aitb_attn_bwd": (_i, [_vp, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp]),
    "aitb_ait_saved_bytes": (_sz, [_i, _i]),
    "aitb_ait_backward_workspace_bytes": (_sz, [_i, _i]),
    "aitb_ait_forward_train": (_i, [C.POINTER(HeadWeights), _vp, _vp, _i, _i, _vp, _vp, _sz, _vp]),
    "aitb_ait_backward": (_i, [C.POINTER(HeadWeights), _vp, _i, _i, _vp, _sz, C.POINTER(AITGrads), _vp, _vp, _vp, _sz,
                               _vp]),
    "aitb_ait_backward_tm": (_i, [C.POINTER(HeadWeights), _vp, _i, _i, _vp, _sz, C.POINTER(AITGrads), _vp, _vp, _vp, _sz,
                                  _vp]),
    "aitb_ait_saved_offset": (_sz, [_i, _i, _i]),
    "aitb_dropout_mask": (_i, [C.c_float, C.c_ulonglong, _i, _i, _vp, _vp]),
    "aitb_attn_dropout_mask": (_i, [C.c_float, C.c_ulonglong, _i, _i, _vp, _vp]),
    "aitb_anchor_target_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "aitb_anchor_target_assign": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _f, _f, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "aitb_anchor_target_finish": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp]),
    "aitb_anchor_target_subsample_device": (_i, [_vp, _vp, _i, _i, _i, _i, C.c_uint64, _vp]),
    "aitb_proposal_target_picks_device": (_i, [_vp, _i, _i, _i, C.c_uint64, _vp, _vp, _vp, _vp]),
    "aitb_proposal_target_assign": (_i, [_vp, _vp, _i, _i, _i, _f, _f, _f, _vp, _vp, _vp, _vp, _vp]),
    "aitb_proposal_target_sample": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _i, C.POINTER(C.c_float),
                                         C.POINTER(C.c_float), C.POINTER(C.c_float), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "aitb_fc_ln": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _f, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "aitb_heads_forward_train": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "aitb_heads_backward_workspace_bytes": (_sz, [_i, _i]),
    "aitb_heads_backward": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "aitb_mean_pool_backward": (_i, [_vp, _i, _vp, _vp]),
    "aitb_relu_bwd": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "aitb_im2col3x3": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "aitb_map_subsample": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "aitb_map_upsample": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "aitb_map_pad": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "aitb_im2col3x3_grouped": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "aitb_sk_combine": (_i, [_vp, _vp, _vp, _sz, _i, _vp]),
    "aitb_sk_combine_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "aitb_rpn_loss": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp]),
    "aitb_rcnn_loss": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
}

_lib = None
_device_checked = False


def load(check_device=True):
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib, _device_checked
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libaitb200.so is not built (%s). Run `python -m ait_b200.build`; there is no "
                "CPU/PyTorch fallback for this path." % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    if check_device and not _device_checked:
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("ait_b200 needs a CUDA device (sm_100a); none is visible")
        if _lib.aitb_check_device() != 0:
            raise RuntimeError(_lib.aitb_last_error().decode())
        _device_checked = True
    return _lib


def check(status):
    if status != 0:
        raise RuntimeError("libaitb200: " + _lib.aitb_last_error().decode())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _first_cuda_device(objs):
    import torch
    for a in objs:
        if isinstance(a, torch.Tensor):
            if a.is_cuda:
                return a.device
        elif isinstance(a, (tuple, list)):
            d = _first_cuda_device(a)
            if d is not None:
                return d
    return None


def on_tensor_device(fn):
    """Decorator of every Python entry point that hands raw pointers to the library: run it with the CUDA device of its
    first CUDA tensor argument current, so `stream_ptr()` (evaluated inside) is that device's current stream, the output
    allocations land there and the launches go to the device that owns the pointers -- also when the caller's current
    device is another one (`model.to('cuda:1')` without `torch.cuda.set_device`).  No-op when the devices already agree."""
    import functools

    @functools.wraps(fn)
    def guarded(*args, **kwargs):
        import torch
        dev = _first_cuda_device(args) or _first_cuda_device(kwargs.values())
        if dev is None or dev.index is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return guarded


def guard_module_functions(namespace, module_name, skip=()):
    """Apply `on_tensor_device` to every public function defined in a module (called at the bottom of ops.py)."""
    import types
    for name, obj in list(namespace.items()):
        if isinstance(obj, types.FunctionType) and obj.__module__ == module_name and not name.startswith("_") \
                and name not in skip:
            namespace[name] = on_tensor_device(obj)


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return C.c_void_p(0 if t is None else t.data_ptr())


def mode_name(compute_dtype):
    """torch.float32 | "fp32" -> "fp32" (split, fp32-class); "tf32"; torch.bfloat16 | "bf16" -> "bf16"."""
    import torch
    if isinstance(compute_dtype, str):
        if compute_dtype not in MODES:
            raise RuntimeError("ait_b200: compute_dtype must be one of %s, torch.float32 or torch.bfloat16; got %r"
                               % (sorted(MODES), compute_dtype))
        return compute_dtype
    if compute_dtype == torch.float32:
        return "fp32"
    if compute_dtype == torch.bfloat16:
        return "bf16"
    raise RuntimeError("ait_b200 supports float32 (fp32-class split or tf32 tensor-core math) and bfloat16 only, "
                       "got %s (the reference's fp64 dispatch is not provided)" % (compute_dtype,))


def storage_dtype(mode):
    import torch
    return torch.float32 if mode == "tf32" else torch.bfloat16


def dtype_enum(torch_dtype):
    import torch
    if torch_dtype == torch.float32:
        return AITB_F32
    if torch_dtype == torch.bfloat16:
        return AITB_BF16
    raise RuntimeError("ait_b200 supports float32 (tf32 tensor cores) and bfloat16 only, got %s "
                       "(the reference's fp64 dispatch is not provided)" % torch_dtype)
