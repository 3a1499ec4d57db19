"""ctypes binding of ``libb2pose.so`` (the C ABI declared in ``include/b2pose.h``).

There is no CPU or PyTorch fallback: if the library is missing, or an entry point reports an
error, a ``RuntimeError`` is raised.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb2pose.so")

F32, BF16 = 0, 1
CONV_PARTIAL, CONV_X_PREMASKED, CONV_DY_PRESCALED, CONV_FORCE_FFMA, CONV_DX_ACCUMULATE, CONV_BN_TOTALS, CONV_W_PREPARED = 1, 2, 4, 8, 16, 32, 64
CONV_WS_HAS_COL = 128
CONV_X_CONCAT = 256
ABI_VERSION = 5
BN_PARTS = 320
MIMIC_PARTS = 64


class ConvDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("N", "H", "W", "C", "K", "R", "S", "stride", "pad", "dil", "Ho", "Wo", "dtype", "flags")]


_p, _i, _l, _f = C.c_void_p, C.c_int32, C.c_int64, C.c_float
_D = C.POINTER(ConvDesc)

# name -> argtypes   (return type is int unless listed in _RESTYPE)
SIGNATURES = {
    "b2_abi_version": [],
    "b2_last_error": [],
    "b2_set_sm_reserve": [_i],
    "b2_conv_uses_tensor_cores": [_D, _i],
    "b2_conv_workspace_bytes": [_D, _i],
    "b2_pconv_fprop": [_D, _p, _p, _p, _p, _p, _p, _p, _p, _p, C.c_size_t, _p],
    "b2_pconv_dgrad": [_D, _p, _p, _p, _p, _p, _p, C.c_size_t, _p],
    "b2_pconv_dgrad_filter_bytes": [_D],
    "b2_pconv_dgrad_filter": [_D, _p, _p, C.c_size_t, _p],
    "b2_pconv_wgrad": [_D, _p, _p, _p, _p, _p, _p, C.c_size_t, _p],
    "b2_pconv_mask_update": [_D, _p, _p, _p, _p],
    "b2_scale_rows": [_p, _p, _p, _l, _i, _i, _p],
    "b2_col_sum": [_p, _p, _p, _l, _i, _i, _p],
    "b2_veil_from_depth": [_p, _p, _l, _i, _p],
    "b2_nchw_to_nhwc": [_p, _p, _i, _i, _i, _i, _i, _p],
    "b2_nhwc_to_nchw": [_p, _p, _i, _i, _i, _i, _i, _p],
    "b2_cast_f32_to_bf16": [_p, _p, _l, _p],
    "b2_cast_bf16_to_f32": [_p, _p, _l, _p],
    "b2_bn_stats": [_p, _l, _i, _i, _p, _p],
    "b2_bn_finalize": [_p, _l, _i, _p, _p, _f, _f, _i, _p, _p, _p],
    "b2_bn_apply": [_p, _p, _p, _p, _p, _p, _p, _i, _p, _l, _i, _i, _p],
    "b2_bn_bwd_reduce": [_p, _p, _p, _p, _p, _p, _p, _p, _i, _p, _l, _i, _i, _p],
    "b2_bn_bwd_finalize": [_p, _i, _p, _p, _p, _p],
    "b2_bn_bwd_apply": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _p, _p, _l, _i, _i, _p],
    "b2_bn_totals_supported": [_i, _i],
    "b2_bn_stats_totals": [_p, _l, _i, _i, _p, _p],
    "b2_bn_apply_totals": [_p, _p, _l, _p, _p, _f, _f, _i, _p, _p, _p, _p, _i, _p, _p, _p, _p, _i, _i, _p],
    "b2_bn_bwd_reduce_totals": [_p, _p, _p, _p, _p, _p, _p, _p, _i, _p, _p, _l, _i, _i, _p],
    "b2_bn_bwd_apply_totals": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _p, _p, _p, _p, _p, _l, _i, _i, _p],
    "b2_mimic_loss_fwd": [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p],
    "b2_mimic_loss_bwd": [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _p],
    "b2_attention_map": [_p, _i, _i, _i, _i, _p, _p],
    "b2_pose_metrics": [_p, _p, _p, _p, _p, _i, _i, _f, _f, _f, _p, _p],
    "b2_remap_normalize_rgb": [_p, _i, _i, _i, _p, _i, C.POINTER(_f), C.POINTER(_f), _p, _p],
    "b2_remap_enhance_depth": [_p, _i, _i, _i, _p, _i, _p, _f, _i, _i, _p, _p],
    "b2_project_points": [_p, _i, C.POINTER(_f), C.POINTER(_f), C.POINTER(_f), C.POINTER(_f), _p, _p],
    "b2_maxpool3x3s2_fwd": [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    "b2_maxpool3x3s2_bwd": [_p, _p, _p, _i, _i, _i, _i, _i, _p],
    "b2_head_fwd": [_p, _i, _i, _i, _i, _i, _i, _i, _f, _p, _p, _p, _p],
    "b2_head_bwd": [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _f, _p, _p],
    "b2_heatmap_softmax": [_p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p],
    "b2_heatmap_decode": [_p, _i, _i, _i, _i, _i, _f, _p, _p],
    "b2_heatmap_decode_bwd": [_p, _i, _i, _i, _i, _i, _f, _p, _p],
    "b2_heatmap_softmax_bwd": [_p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _p],
    "b2_pose_loss": [_p, _p, _p, _i, _i, _i, _f, _i, _p, _p, _p, _p],
    "b2_unproject_depth": [_p, _p, _i, _i, _i, C.POINTER(_f), C.POINTER(_f), _p],
    "b2_grad_sumsq": [_p, _l, _p, _p],
    "b2_adam_step": [_p, _p, _p, _p, _p, _l, _f, _f, _f, _f, _f, _i, _p, _f, _f, _p, _p],
}
_RESTYPE = {"b2_last_error": C.c_char_p, "b2_conv_workspace_bytes": C.c_size_t,
            "b2_pconv_dgrad_filter_bytes": C.c_size_t}

_lib = None
launches = 0      # number of kernel-launching entry-point calls made through this binding


def lib():
    """The loaded library (loads on first use; raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libb2pose.so is not built (%s); run `python __graft_entry__.py` -- "
                               "there is no CPU fallback" % LIB_PATH)
        h = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(h, name)
            fn.argtypes = args
            fn.restype = _RESTYPE.get(name, C.c_int)
        if h.b2_abi_version() != ABI_VERSION:
            raise RuntimeError("libb2pose.so ABI %d != binding ABI %d" % (h.b2_abi_version(), ABI_VERSION))
        _lib = h
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().b2_last_error()
        raise RuntimeError("libb2pose %s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))


_SYNC_DEBUG = os.environ.get("B2POSE_SYNC") == "1"     # debugging aid: localise asynchronous faults


def call(name, *args):
    """Call an int-returning entry point and raise on a non-zero status."""
    global launches
    launches += 1
    check(getattr(lib(), name)(*args), name)
    if _SYNC_DEBUG:
        try:
            torch.cuda.synchronize()
        except Exception as e:      # noqa: BLE001
            desc = ""
            for a in args:
                d = getattr(a, "_obj", None)
                if isinstance(d, ConvDesc):
                    desc = " desc=" + str({f: getattr(d, f) for f, _ in ConvDesc._fields_})
            raise RuntimeError("libb2pose %s faulted%s: %s" % (name, desc, e)) from e


class TensorPair:
    """Two NHWC tensors standing for their channel concatenation (B2_CONV_X_CONCAT): passed to the C ABI as an array
    of two device pointers."""

    def __init__(self, a, b):
        assert a.shape == b.shape and a.dtype == b.dtype and a.device == b.device
        self.a, self.b = a, b
        self.device, self.dtype = a.device, a.dtype
        self._ptrs = (C.c_void_p * 2)(a.data_ptr(), b.data_ptr())


def ptr(t):
    """Device pointer of a tensor (None -> NULL; a TensorPair -> host array of its two device pointers)."""
    if isinstance(t, TensorPair):
        return t._ptrs
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def dt(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError("libb2pose computes in float32 or bfloat16, got %s" % t.dtype)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("b2pose runs on CUDA tensors only (got a %s tensor); there is no CPU fallback"
                               % t.device)
