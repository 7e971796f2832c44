"""ctypes binding of libmanipose_sm100.so (include/manipose_sm100.h).

The library is the product: if it is missing or a call fails, this module raises — there is no CPU or
PyTorch fallback (BASELINE.json north_star).  Every wrapper passes raw device pointers and the current
CUDA stream; nothing here allocates device memory.
"""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmanipose_sm100.so")

MP_OK = 0
MP_EINVAL, MP_EUNSUPPORTED, MP_EDEVICE, MP_ELAUNCH, MP_EALIGN, MP_EWORKSPACE = -1, -2, -3, -4, -5, -6
MP_DEC_EXACT, MP_DEC_FAST = 0, 1
MP_TERM_WTA, MP_TERM_BCE, MP_TERM_VEL, MP_TERM_SMOOTH, MP_TERM_TOTAL, MP_LOSS_NTERMS = 0, 1, 2, 3, 4, 8
MP_AGG_WEIGHTED_AVE, MP_AGG_BEST_SCORE, MP_AGG_ORACLE = 0, 1, 2
MP_EPI_BIAS, MP_EPI_GELU, MP_EPI_RESIDUAL, MP_EPI_ACCUMULATE, MP_EPI_BIAS_F32 = 0, 1, 2, 3, 4
MP_ATTN_SPATIAL, MP_ATTN_TEMPORAL = 0, 1
MP_DTYPE_BF16, MP_DTYPE_FP16 = 0, 1
MP_ERR_L2, MP_ERR_SQ, MP_ERR_ABS, MP_ERR_DIFF = 0, 1, 2, 3

P, I64, F, I = c_void_p, c_int64, c_float, c_int

# name -> (restype, argtypes); mirrors include/manipose_sm100.h declaration by declaration
SIGNATURES = {
    "mp_abi_version": (I, []),
    "mp_last_error": (c_char_p, []),
    "mp_device_check": (I, []),
    "mp_set_sm_limit": (I, [I]),
    "mp_set_skeleton": (I, [I, ctypes.POINTER(c_int32), ctypes.POINTER(c_float)]),
    "mp_decoder_fwd": (I, [P, P, P, P, P, P, I64, I64, I64, I, I, P]),
    "mp_decoder_bwd_workspace_bytes": (c_size_t, [I64, I64, I64]),
    "mp_decoder_bwd": (I, [P, P, P, P, P, P, I64, I64, I64, I, P, c_size_t, P]),
    "mp_softmax_hyp_fwd": (I, [P, P, I64, I64, I64, P]),
    "mp_softmax_hyp_bwd": (I, [P, P, P, I64, I64, I64, P]),
    "mp_wta_fwd": (I, [P, P, P, I, P, P, P, I64, I64, I64, P]),
    "mp_loss_workspace_bytes": (c_size_t, [I64, I64, I64]),
    "mp_loss_fwd": (I, [P, P, P, P, I, F, F, F, P, P, P, I64, I64, I64, P, c_size_t, P]),
    "mp_loss_bwd": (I, [P, P, P, P, P, I, F, F, F, P, P, P, P, I64, I64, I64, P]),
    "mp_aggregate": (I, [P, P, P, I, P, P, P, I64, I64, I64, P]),
    "mp_aggregate_tta": (I, [P, P, I, P, I64, I64, I64, P]),
    "mp_mpjpe_workspace_bytes": (c_size_t, [I64]),
    "mp_mpjpe": (I, [P, P, I64, P, P, c_size_t, P]),
    "mp_linear": (I, [P, P, P, P, P, I64, I64, I64, I, I, P]),
    "mp_linear_gelu2": (I, [P, P, P, P, P, I64, I64, I64, I, P]),
    "mp_linear_ln": (I, [P, P, P, P, P, P, P, P, F, P, I64, I64, P, P, F, P, P, I64, I64, I64, I, P]),
    "mp_mlp_ln": (I, [P, P, P, P, P, P, P, P, P, P, F, P, I64, I64, P, P, F, I64, I64, I64, I, P]),
    "mp_layernorm": (I, [P, P, P, P, P, F, P, I64, I64, P, P, F, I64, I, I, P]),
    "mp_embed_joints": (I, [P, P, P, P, P, P, F, P, P, I64, I, I, I, P]),
    "mp_embed_segments": (I, [P, P, P, P, P, P, F, P, P, I64, I, I, I, I, P]),
    "mp_attention": (I, [P, P, I64, I64, I, I, I, I, I, P]),
    "mp_heads_fwd": (I, [P, P, P, F, P, P, P, P, P, P, P, P, I64, I64, I, I, I, P]),
    "mp_heads_fwd16": (I, [P, P, P, P, P, P, P, P, c_size_t, I64, I64, I, I, I, I, I, P]),
    "mp_bones_head": (I, [P, P, P, F, P, P, P, P, P, I64, I64, I, I, P, c_size_t, P]),
    "mp_cast_f32_to_16": (I, [P, P, I64, I, P]),
    "mp_gather_windows": (I, [P, P, P, P, P, P, P, P, P, I64, I64, I, I, P]),
    "mp_pck_auc_workspace_bytes": (c_size_t, []),
    "mp_pck_auc": (I, [P, P, I64, F, P, P, c_size_t, P]),
    "mp_p_mpjpe_workspace_bytes": (c_size_t, [I64]),
    "mp_p_mpjpe": (I, [P, P, I64, P, P, c_size_t, P]),
    "mp_pose_consistency_workspace_bytes": (c_size_t, [I64, I64]),
    "mp_pose_consistency": (I, [P, I64, I64, P, P, P, P, P, P, c_size_t, P]),
    "mp_point_errors_workspace_bytes": (c_size_t, [I64, I]),
    "mp_point_errors": (I, [P, P, I64, I, I, F, P, P, P, c_size_t, P]),
    "mp_layernorm_bwd": (I, [P, P, F, P, I, P, P, P, P, P, P, P, I64, I, I, P]),
    "mp_gelu_fwd": (I, [P, P, I64, I, P]),
    "mp_gelu_bwd": (I, [P, P, P, I64, I, P]),
    "mp_gelu_bwd_colsum": (I, [P, P, P, P, I64, I64, I, P]),
    "mp_attention_bwd": (I, [P, P, P, P, P, I64, I64, I, I, I, I, I, P]),
    "mp_wgrad": (I, [P, P, P, I64, I64, I64, I, P]),
    "mp_colsum16": (I, [P, P, I64, I64, I, P]),
    "mp_heads_fold": (I, [P, I, I, I, I, P, P, P, P, P, I, P]),
    "mp_heads_bwd_pack": (I, [P, P, P, P, P, P, P, I64, I64, I, I, I, I, P]),
    "mp_heads_unfold": (I, [P, P, P, P, I, I, I, P]),
    "mp_refresh_shadows": (I, [P, I, I, I, P]),
    "mp_transpose16": (I, [P, P, P, I64, I64, I64, I, P]),
    "mp_group_rowsum": (I, [P, P, I64, I, I64, I64, P]),
    "mp_small_wgrad": (I, [P, P, P, P, I64, I, I, P]),
    "mp_residual_rowscale": (I, [P, P, P, P, I64, I, I, P]),
    "mp_cast_rowscale": (I, [P, P, P, I64, I, I, P]),
    "mp_adam_step": (I, [P, P, P, P, I64, F, F, F, F, F, I64, P, P, F, P]),
}

_lib = None


class ManiposeLibraryError(RuntimeError):
    """The sm_100a extension is missing, or a launch failed (MP_EDEVICE / MP_ELAUNCH / MP_EWORKSPACE / MP_EALIGN)."""


def load():
    """Loads the in-tree shared library (built by ``python -m manipose_b200._build`` / ``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ManiposeLibraryError(
            f"{LIB_PATH} not found: build it with `python -m manipose_b200._build` (nvcc, sm_100a). "
            "manipose_b200 has no CPU / PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    msg = load().mp_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, what: str = ""):
    """Maps a negative MP_E* code to the exception type the reference would raise for the same mistake:
    MP_EINVAL -> AssertionError / ValueError territory (we raise ValueError unless the message is one of the
    reference's assert messages), MP_EUNSUPPORTED -> NotImplementedError, everything else -> ManiposeLibraryError."""
    if rc == MP_OK:
        return
    msg = last_error() or what
    if rc == MP_EINVAL:
        if msg.startswith(("Unsupported rotations", "Scores required", "Ground-truth required")):
            raise AssertionError(msg)
        raise ValueError(msg)
    if rc == MP_EUNSUPPORTED:
        raise NotImplementedError(msg)
    raise ManiposeLibraryError(f"{what or 'libmanipose_sm100'} failed with code {rc}: {msg}")


def stream_ptr():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else c_void_p(t.data_ptr())
