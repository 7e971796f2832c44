"""Tensor-level wrappers over the C ABI (include/manipose_sm100.h).

PyTorch is plumbing here: it owns device memory and streams.  All arithmetic happens in libmanipose_sm100.so;
a CPU tensor or a missing library raises (no fallback).
"""
from typing import Optional, Tuple

import torch

from . import _lib as L

J = 17
BONES = 16

# Kernel launches issued through this module (bench.py reports it as gpu_launches) and an optional sampling hook:
# when GEMM_TIMING is a list, linear() / linear_ln() bracket their launch with CUDA events and append
# (start, end, algorithmic flops, algorithmic bytes, kernel family).
LAUNCHES = 0
GEMM_TIMING = None


def _count(n=1):
    global LAUNCHES
    LAUNCHES += n


class nvtx:
    """NVTX range around a stage of the path (trunk / heads / decoder / loss ...): the ranges show up in Nsight Systems / Compute
    timelines of the drivers.  A push / pop pair costs about a microsecond on the host and nothing on the device."""

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        torch.cuda.nvtx.range_push(self.name)
        return self

    def __exit__(self, *exc):
        torch.cuda.nvtx.range_pop()
        return False


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise L.ManiposeLibraryError(
                "manipose_b200 runs on sm_100a only: got a CPU tensor (there is no CPU fallback; "
                "move the model and its inputs to a B200)")


def _f32(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


_skeleton_checked = set()


def set_skeleton(parents, t_pose_operators) -> None:
    """Validates a skeleton (parents, t_pose_operators) against the tree the kernels are specialised for
    (hpe/mh_so3_hpe/data/skeleton.py; h36m_lifting.py:40-57).  Raises NotImplementedError for another tree."""
    import ctypes
    par = [int(p) for p in parents]
    ops = [float(v) for row in t_pose_operators for v in (row.tolist() if hasattr(row, "tolist") else row)]
    key = (tuple(par), tuple(ops))
    if key in _skeleton_checked:
        return
    n = len(par)
    if len(ops) != 3 * n:
        raise ValueError(f"t_pose_operators must have {n} rows of 3, got {len(ops)} values")
    rc = L.load().mp_set_skeleton(n, (ctypes.c_int32 * n)(*par), (ctypes.c_float * (3 * n))(*ops))
    L.check(rc, "mp_set_skeleton")
    _skeleton_checked.add(key)


# ------------------------------------------------------------------------------------------------ decoder
def decoder_fwd(rot6d: torch.Tensor, bone_len: torch.Tensor, root: Optional[torch.Tensor], logits: Optional[torch.Tensor],
                n_clips: int, n_hyp: int, n_frames: int, rot_rep_dim: int = 6, exact: bool = True):
    _need_cuda(rot6d, bone_len, root, logits)
    rot6d, bone_len, root, logits = _f32(rot6d), _f32(bone_len), _f32(root), _f32(logits)
    n_poses = n_clips * n_hyp * n_frames
    poses = torch.empty((n_poses, J, 3), dtype=torch.float32, device=rot6d.device)
    scores = torch.empty_like(logits) if logits is not None else None
    rc = L.load().mp_decoder_fwd(L.ptr(rot6d), L.ptr(bone_len), L.ptr(root), L.ptr(logits), L.ptr(poses), L.ptr(scores),
                                 n_clips, n_hyp, n_frames, rot_rep_dim, L.MP_DEC_EXACT if exact else L.MP_DEC_FAST, L.stream_ptr())
    L.check(rc, "mp_decoder_fwd")
    _count()
    return poses, scores


class _DecoderFn(torch.autograd.Function):
    """poses = FK(GramSchmidt(rot6d), bone_len, root)   (PoseDecoder.forward, pose_decoder.py:32-55)."""

    @staticmethod
    def forward(ctx, rot6d, bone_len, root, n_clips, n_hyp, n_frames, rot_rep_dim, exact):
        rot6d, bone_len = _f32(rot6d), _f32(bone_len)
        poses, _ = decoder_fwd(rot6d, bone_len, root, None, n_clips, n_hyp, n_frames, rot_rep_dim, exact)
        ctx.save_for_backward(rot6d, bone_len)
        ctx.dims = (n_clips, n_hyp, n_frames, rot_rep_dim)
        ctx.root_grad = root is not None and root.requires_grad
        ctx.bone_shape = bone_len.shape
        return poses

    @staticmethod
    def backward(ctx, g):
        rot6d, bone_len = ctx.saved_tensors
        n_clips, n_hyp, n_frames, d = ctx.dims
        g = _f32(g)
        g_rot = torch.empty_like(rot6d)
        g_bone = torch.empty(ctx.bone_shape, dtype=torch.float32, device=rot6d.device)
        g_root = torch.empty((n_clips * n_hyp * n_frames, 3), dtype=torch.float32, device=rot6d.device) if ctx.root_grad else None
        wsb = L.load().mp_decoder_bwd_workspace_bytes(n_clips, n_hyp, n_frames)
        ws = torch.empty(wsb, dtype=torch.uint8, device=rot6d.device)
        rc = L.load().mp_decoder_bwd(L.ptr(rot6d), L.ptr(bone_len), L.ptr(g), L.ptr(g_rot), L.ptr(g_bone), L.ptr(g_root),
                                     n_clips, n_hyp, n_frames, d, L.ptr(ws), wsb, L.stream_ptr())
        L.check(rc, "mp_decoder_bwd")
        _count(2)
        return g_rot, g_bone, g_root, None, None, None, None, None


def decode(rot6d, bone_len, root, n_clips, n_hyp, n_frames, rot_rep_dim=6, exact=True):
    """Differentiable decoder: rot6d [n_clips*n_hyp*n_frames, 17, D], bone_len [n_clips, 16(,1)] -> poses [N,17,3]."""
    _need_cuda(rot6d, bone_len, root)
    return _DecoderFn.apply(rot6d, bone_len, root, n_clips, n_hyp, n_frames, rot_rep_dim, exact)


class _SoftmaxHypFn(torch.autograd.Function):
    """softmax over the hypothesis dim of logits [B,K,T(,1)] (rmcl_manifold_mix_ste.py:262)."""

    @staticmethod
    def forward(ctx, logits):
        lg = _f32(logits)
        b, k, t = lg.shape[:3]
        scores = torch.empty_like(lg)
        L.check(L.load().mp_softmax_hyp_fwd(L.ptr(lg), L.ptr(scores), b, k, t, L.stream_ptr()), "mp_softmax_hyp_fwd")
        ctx.save_for_backward(scores)
        ctx.dims = (b, k, t)
        return scores

    @staticmethod
    def backward(ctx, g):
        (scores,) = ctx.saved_tensors
        b, k, t = ctx.dims
        g = _f32(g)
        out = torch.empty_like(scores)
        L.check(L.load().mp_softmax_hyp_bwd(L.ptr(scores), L.ptr(g), L.ptr(out), b, k, t, L.stream_ptr()), "mp_softmax_hyp_bwd")
        return out


def softmax_hyp(logits: torch.Tensor) -> torch.Tensor:
    _need_cuda(logits)
    return _SoftmaxHypFn.apply(logits)


# ------------------------------------------------------------------------------------------------ losses / metrics
_weights_cache = {}


def _weights_dev(weights, device):
    """Per-joint loss weights on the device.  Host tensors (STANDARD_H36M_WEIGHTS is one) are uploaded once and cached by value, so
    the loss can run inside a CUDA-graph capture (no pageable host copy in the step)."""
    if weights is None:
        return None
    if weights.is_cuda:
        return _f32(weights)
    key = (str(device), tuple(float(v) for v in weights.tolist()))
    w = _weights_cache.get(key)
    if w is None:
        w = _f32(weights.to(device))
        _weights_cache[key] = w
    return w


def loss_workspace(device, b, k, t):
    n = L.load().mp_loss_workspace_bytes(b, k, t)
    return torch.empty(n, dtype=torch.uint8, device=device), n


class _LossTermsFn(torch.autograd.Function):
    """terms[8] = [wta mean, bce mean, velocity mean, smoothness mean, weighted total, 0, 0, 0], wta_val [B,T],
    wta_idx [B,T] int64 — all of hpe/main_h36m_lifting.py:129-169 in one pass."""

    @staticmethod
    def forward(ctx, hyp, scores, y, weights, squared, beta, vel_w, smooth_w):
        hyp, scores, y = _f32(hyp), _f32(scores), _f32(y)
        b, k, t = hyp.shape[:3]
        dev = hyp.device
        terms = torch.empty(L.MP_LOSS_NTERMS, dtype=torch.float32, device=dev)
        wta_val = torch.empty((b, t), dtype=torch.float32, device=dev)
        wta_idx = torch.empty((b, t), dtype=torch.int64, device=dev)
        ws, nbytes = loss_workspace(dev, b, k, t)
        rc = L.load().mp_loss_fwd(L.ptr(hyp), L.ptr(scores), L.ptr(y), L.ptr(weights), int(squared), float(beta), float(vel_w),
                                  float(smooth_w), L.ptr(terms), L.ptr(wta_val), L.ptr(wta_idx), b, k, t, L.ptr(ws), nbytes,
                                  L.stream_ptr())
        L.check(rc, "mp_loss_fwd")
        _count(2)
        ctx.save_for_backward(hyp, scores, y, weights, wta_idx)
        ctx.cfg = (int(squared), float(beta), float(vel_w), float(smooth_w), b, k, t)
        ctx.mark_non_differentiable(wta_idx)
        return terms, wta_val, wta_idx

    @staticmethod
    def backward(ctx, g_terms, g_wta_val, _g_idx):
        hyp, scores, y, weights, wta_idx = ctx.saved_tensors
        squared, beta, vel_w, smooth_w, b, k, t = ctx.cfg
        dev = hyp.device
        g_terms = _f32(g_terms) if g_terms is not None else torch.zeros(L.MP_LOSS_NTERMS, dtype=torch.float32, device=dev)
        g_wta_val = _f32(g_wta_val) if g_wta_val is not None else None
        g_hyp = torch.empty_like(hyp)
        g_scores = torch.empty_like(scores) if (scores is not None and ctx.needs_input_grad[1]) else None
        rc = L.load().mp_loss_bwd(L.ptr(hyp), L.ptr(scores), L.ptr(y), L.ptr(weights), L.ptr(wta_idx), squared, beta, vel_w,
                                  smooth_w, L.ptr(g_terms), L.ptr(g_wta_val), L.ptr(g_hyp), L.ptr(g_scores), b, k, t,
                                  L.stream_ptr())
        L.check(rc, "mp_loss_bwd")
        _count()
        return g_hyp, g_scores, None, None, None, None, None, None


def loss_terms(hyp, scores, y, weights=None, squared=False, beta=0.0, vel_w=0.0, smooth_w=0.0):
    """hyp [B,K,T,17,3], scores [B,K,T(,1)] or None, y [B,T,17,3] -> (terms[8], wta_val[B,T], wta_idx[B,T])."""
    _need_cuda(hyp, scores, y)
    if hyp.dim() != 5 or hyp.shape[-2:] != (J, 3) or y.shape != (hyp.shape[0], hyp.shape[2], J, 3):
        raise ValueError(f"expected hypothesis [B,K,T,17,3] and y [B,T,17,3], got {tuple(hyp.shape)} and {tuple(y.shape)}")
    w = _weights_dev(weights, hyp.device)
    if w is not None and w.shape[0] != J:
        raise AssertionError("weights.shape[0] == target.shape[-2]")
    return _LossTermsFn.apply(hyp, scores, y, w, bool(squared), beta, vel_w, smooth_w)


def wta_fwd(hyp, y, weights=None, squared=False, per_hyp=False):
    """No-grad WTA: (values [B,T], indices [B,T] int64[, per-hypothesis errors [B,K,T]])."""
    _need_cuda(hyp, y)
    hyp, y = _f32(hyp), _f32(y)
    b, k, t = hyp.shape[:3]
    w = _weights_dev(weights, hyp.device)
    if w is not None and w.shape[0] != J:
        raise AssertionError("weights.shape[0] == target.shape[-2]")
    val = torch.empty((b, t), dtype=torch.float32, device=hyp.device)
    idx = torch.empty((b, t), dtype=torch.int64, device=hyp.device)
    ph = torch.empty((b, k, t), dtype=torch.float32, device=hyp.device) if per_hyp else None
    rc = L.load().mp_wta_fwd(L.ptr(hyp), L.ptr(y), L.ptr(w), int(squared), L.ptr(val), L.ptr(idx), L.ptr(ph), b, k, t, L.stream_ptr())
    L.check(rc, "mp_wta_fwd")
    _count()
    return (val, idx, ph) if per_hyp else (val, idx)


def aggregate(hyp, scores=None, y=None, mode=L.MP_AGG_WEIGHTED_AVE):
    """RMCLManifoldMixSTE.aggregate on device: returns (pose [B,T,17,3], val [B,T] or None, idx [B,T] or None)."""
    _need_cuda(hyp, scores, y)
    hyp, scores, y = _f32(hyp), _f32(scores), _f32(y)
    b, k, t = hyp.shape[:3]
    dev = hyp.device
    pose = torch.empty((b, t, J, 3), dtype=torch.float32, device=dev)
    val = torch.empty((b, t), dtype=torch.float32, device=dev) if mode == L.MP_AGG_ORACLE else None
    idx = torch.empty((b, t), dtype=torch.int64, device=dev) if mode != L.MP_AGG_WEIGHTED_AVE else None
    rc = L.load().mp_aggregate(L.ptr(hyp), L.ptr(scores), L.ptr(y), mode, L.ptr(pose), L.ptr(val), L.ptr(idx), b, k, t, L.stream_ptr())
    L.check(rc, "mp_aggregate")
    _count(1 if mode == L.MP_AGG_WEIGHTED_AVE else 2)
    return pose, val, idx


def aggregate_tta(hyp2, scores2, mode=L.MP_AGG_WEIGHTED_AVE):
    """hyp2 [2B,K,T,17,3], scores2 [2B,K,T(,1)] (clips [B,2B) = forward of the flipped input) -> TTA prediction [B,T,17,3]."""
    _need_cuda(hyp2, scores2)
    hyp2, scores2 = _f32(hyp2), _f32(scores2)
    b2, k, t = hyp2.shape[:3]
    if b2 % 2 != 0:
        raise ValueError("aggregate_tta expects the original and the flipped half stacked on dim 0")
    out = torch.empty((b2 // 2, t, J, 3), dtype=torch.float32, device=hyp2.device)
    L.check(L.load().mp_aggregate_tta(L.ptr(hyp2), L.ptr(scores2), mode, L.ptr(out), b2 // 2, k, t, L.stream_ptr()), "mp_aggregate_tta")
    _count()
    return out


def mpjpe(pred, gt):
    """(sum, mean) of ||gt - pred||_2 over all 3-D points, as a 2-element fp32 device tensor."""
    _need_cuda(pred, gt)
    pred, gt = _f32(pred), _f32(gt)
    if pred.shape[-1] != 3 or gt.shape[-1] != 3 or pred.numel() != gt.numel():
        raise AssertionError("batch_imp.shape[-1] == batch_gt.shape[-1] == 3")
    n = pred.numel() // 3
    out = torch.empty(2, dtype=torch.float32, device=pred.device)
    nbytes = L.load().mp_mpjpe_workspace_bytes(n)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=pred.device)
    rc = L.load().mp_mpjpe(L.ptr(pred), L.ptr(gt), n, L.ptr(out), L.ptr(ws), nbytes, L.stream_ptr())
    L.check(rc, "mp_mpjpe")
    _count(2)
    return out


def p_mpjpe(pred, gt):
    """(sum, mean) of the joint distances after per-frame Procrustes alignment, as a 2-element fp32 device tensor."""
    _need_cuda(pred, gt)
    pred, gt = _f32(pred), _f32(gt)
    if pred.shape != gt.shape or pred.shape[-2:] != (J, 3):
        raise AssertionError("predicted.shape == target.shape == [..., 17, 3]")
    n = pred.numel() // (J * 3)
    out = torch.empty(2, dtype=torch.float32, device=pred.device)
    nbytes = L.load().mp_p_mpjpe_workspace_bytes(n)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=pred.device)
    L.check(L.load().mp_p_mpjpe(L.ptr(pred), L.ptr(gt), n, L.ptr(out), L.ptr(ws), nbytes, L.stream_ptr()), "mp_p_mpjpe")
    _count(2)
    return out


def pck_auc(pred, gt, threshold=150.0):
    """(PCK at threshold, AUC over linspace(0, 150, 31)) in percent, as a 2-element fp32 device tensor."""
    _need_cuda(pred, gt)
    pred, gt = _f32(pred), _f32(gt)
    if pred.shape != gt.shape or pred.shape[-1] != 3:
        raise AssertionError("pred.shape == gt.shape == [..., 3]")
    out = torch.empty(2, dtype=torch.float32, device=pred.device)
    nbytes = L.load().mp_pck_auc_workspace_bytes()
    ws = torch.empty(nbytes, dtype=torch.uint8, device=pred.device)
    rc = L.load().mp_pck_auc(L.ptr(pred), L.ptr(gt), pred.numel() // 3, float(threshold), L.ptr(out), L.ptr(ws), nbytes, L.stream_ptr())
    L.check(rc, "mp_pck_auc")
    _count(2)
    return out


def point_errors(pred: torch.Tensor, gt: torch.Tensor, mode: int, cols: int = 0, scale: float = 1.0, per_elem: bool = False):
    """Per-point / per-column errors (mean_joint_errors.py:31-141) in one streaming pass.  mode L2 / SQ: pred, gt [..., 3] points;
    ABS / DIFF: scalars.  Returns (per-element errors [n] or None, scale * column sums [cols] or None)."""
    _need_cuda(pred, gt)
    pred, gt = _f32(pred), _f32(gt)
    if pred.numel() != gt.numel():
        raise AssertionError("batch_imp and batch_gt must hold the same number of values")
    vec = 3 if mode in (L.MP_ERR_L2, L.MP_ERR_SQ) else 1
    n = pred.numel() // vec
    dev = pred.device
    elems = torch.empty(n, dtype=torch.float32, device=dev) if per_elem else None
    col = torch.empty(cols, dtype=torch.float32, device=dev) if cols > 0 else None
    c = max(cols, 1)
    nbytes = L.load().mp_point_errors_workspace_bytes(n, c) if cols > 0 else 0
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
    rc = L.load().mp_point_errors(L.ptr(pred), L.ptr(gt), n, c, mode, float(scale), L.ptr(elems), L.ptr(col), L.ptr(ws), nbytes, L.stream_ptr())
    L.check(rc, "mp_point_errors")
    _count(2 if cols > 0 else 1)
    return elems, col


def pose_consistency(poses: torch.Tensor, with_bone_lengths: bool = False):
    """poses [B, L, 17, 3] -> (seg_mean [B,16], seg_var [B,16] (unbiased, over time), sym_abs [B,6], sym_sq [B,6], bone_len [B,16,L] | None)."""
    _need_cuda(poses)
    poses = _f32(poses)
    if poses.dim() != 4 or poses.shape[-2:] != (J, 3):
        raise ValueError(f"expected poses [B, L, 17, 3], got {tuple(poses.shape)}")
    b, l = poses.shape[:2]
    dev = poses.device
    seg_mean = torch.empty((b, BONES), dtype=torch.float32, device=dev)
    seg_var = torch.empty((b, BONES), dtype=torch.float32, device=dev)
    sym_abs = torch.empty((b, 6), dtype=torch.float32, device=dev)
    sym_sq = torch.empty((b, 6), dtype=torch.float32, device=dev)
    bone_len = torch.empty((b, BONES, l), dtype=torch.float32, device=dev) if with_bone_lengths else None
    nbytes = L.load().mp_pose_consistency_workspace_bytes(b, l)
    ws = torch.empty(max(nbytes, 4), dtype=torch.uint8, device=dev)
    rc = L.load().mp_pose_consistency(L.ptr(poses), b, l, L.ptr(seg_mean), L.ptr(seg_var), L.ptr(sym_abs), L.ptr(sym_sq), L.ptr(bone_len),
                                      L.ptr(ws), nbytes, L.stream_ptr())
    L.check(rc, "mp_pose_consistency")
    _count(2)
    return seg_mean, seg_var, sym_abs, sym_sq, bone_len


# ------------------------------------------------------------------------------------------------ backbone pieces
TORCH_DTYPE = {L.MP_DTYPE_BF16: torch.bfloat16, L.MP_DTYPE_FP16: torch.float16}
DTYPE_CODE = {"bf16": L.MP_DTYPE_BF16, "fp16": L.MP_DTYPE_FP16, torch.bfloat16: L.MP_DTYPE_BF16, torch.float16: L.MP_DTYPE_FP16}


def cast16(src: torch.Tensor, dtype: int, dst: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 -> bf16 / fp16 (round to nearest even; fp16 saturates to +-65504)."""
    _need_cuda(src)
    src = _f32(src)
    if dst is None:
        dst = torch.empty(src.shape, dtype=TORCH_DTYPE[dtype], device=src.device)
    rc = L.load().mp_cast_f32_to_16(L.ptr(src), L.ptr(dst), src.numel(), dtype, L.stream_ptr())
    L.check(rc, "mp_cast_f32_to_16")
    _count()
    return dst


def linear(a, w, bias, out, epilogue=L.MP_EPI_BIAS, resid=None):
    """out[M,N] = epilogue(a[M,K] @ w[N,K]^T + bias[N]); a, w 16-bit (same dtype); bias fp32;
    out 16-bit (BIAS / GELU) or fp32 with resid fp32 (RESIDUAL; out may alias resid)."""
    m, k = a.shape
    n = w.shape[0]
    dtype = DTYPE_CODE[a.dtype]
    timing = GEMM_TIMING
    if timing is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = L.load().mp_linear(L.ptr(a), L.ptr(w), L.ptr(bias), L.ptr(resid), L.ptr(out), m, n, k, epilogue, dtype, L.stream_ptr())
    L.check(rc, "mp_linear")
    _count()
    if timing is not None:
        e1.record()
        esz = 4 if epilogue == L.MP_EPI_RESIDUAL else 2
        byts = 2.0 * (m * k + n * k) + esz * m * n * (2 if epilogue == L.MP_EPI_RESIDUAL else 1)
        timing.append((e0, e1, 2.0 * m * n * k, byts, "linear"))
    return out


def linear_gelu2(a, w, bias, u_out, g_out):
    """u_out = a @ w^T + bias, g_out = GELU(a @ w^T + bias) (both 16-bit) in one launch: fc1 of the training forward."""
    m, k = a.shape
    n = w.shape[0]
    rc = L.load().mp_linear_gelu2(L.ptr(a), L.ptr(w), L.ptr(bias), L.ptr(u_out), L.ptr(g_out), m, n, k, DTYPE_CODE[a.dtype], L.stream_ptr())
    L.check(rc, "mp_linear_gelu2")
    _count()


def linear_ln(a, w, bias, resid, x_out, h_out, post=None, post_eps=1e-6, pos=None, pos_div=1, pos_mod=1, ln=None, ln_eps=1e-6, row_scale=None, x_pre=None):
    """x_out = [LN_post](resid + s * (a @ w^T + bias)) [+ pos]; h_out = LN_pre(x_out) as 16-bit; s = row_scale[row] (fp32 [M]) or 1;
    x_pre (with post) also receives the value before LN_post.      N must be 512 (fused epilogue)."""
    m, k = a.shape
    n = w.shape[0]
    pg, pb = post if post is not None else (None, None)
    lg, lb = ln if ln is not None else (None, None)
    timing = GEMM_TIMING
    if timing is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = L.load().mp_linear_ln(L.ptr(a), L.ptr(w), L.ptr(bias), L.ptr(resid), L.ptr(x_out), L.ptr(h_out), L.ptr(pg), L.ptr(pb), post_eps,
                               L.ptr(pos), pos_div, pos_mod, L.ptr(lg), L.ptr(lb), ln_eps, L.ptr(row_scale), L.ptr(x_pre), m, n, k, DTYPE_CODE[a.dtype], L.stream_ptr())
    L.check(rc, "mp_linear_ln")
    _count()
    if timing is not None:
        e1.record()
        byts = 2.0 * (m * k + n * k) + 8.0 * m * n + (2.0 * m * n if ln is not None else 0.0)
        timing.append((e0, e1, 2.0 * m * n * k, byts, "linear_ln"))


def mlp_ln(h_in, w1, b1, w2, b2, resid, x_out, h_out, post=None, post_eps=1e-6, pos=None, pos_div=1, pos_mod=1, ln=None, ln_eps=1e-6):
    """x_out = [LN_post](resid + fc2(gelu(fc1(h_in)))) [+ pos]; h_out = LN_pre(x_out) as 16-bit — the MLP branch of a C = 512 block in one
    launch, the hidden activation never leaves the SM.  h_out may alias h_in."""
    m, c = h_in.shape
    hidden = w1.shape[0]
    pg, pb = post if post is not None else (None, None)
    lg, lb = ln if ln is not None else (None, None)
    timing = GEMM_TIMING
    if timing is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = L.load().mp_mlp_ln(L.ptr(h_in), L.ptr(w1), L.ptr(b1), L.ptr(w2), L.ptr(b2), L.ptr(resid), L.ptr(x_out), L.ptr(h_out), L.ptr(pg),
                            L.ptr(pb), post_eps, L.ptr(pos), pos_div, pos_mod, L.ptr(lg), L.ptr(lb), ln_eps, m, c, hidden, DTYPE_CODE[h_in.dtype],
                            L.stream_ptr())
    L.check(rc, "mp_mlp_ln")
    _count()
    if timing is not None:
        e1.record()
        byts = 2.0 * (m * c + 2 * hidden * c) + (8.0 * m * c if x_out is not None else 4.0 * m * c) + (2.0 * m * c if ln is not None else 0.0)
        timing.append((e0, e1, 4.0 * m * hidden * c, byts, "mlp"))


def layernorm(x_in, x_out, h_out, post=None, post_eps=1e-6, pos=None, pos_div=1, pos_mod=1, ln=None, ln_eps=1e-6, dtype=L.MP_DTYPE_BF16):
    n_tokens, c = x_in.shape
    pg, pb = post if post is not None else (None, None)
    lg, lb = ln if ln is not None else (None, None)
    rc = L.load().mp_layernorm(L.ptr(x_in), L.ptr(x_out), L.ptr(h_out), L.ptr(pg), L.ptr(pb), post_eps, L.ptr(pos), pos_div, pos_mod,
                               L.ptr(lg), L.ptr(lb), ln_eps, n_tokens, c, dtype, L.stream_ptr())
    L.check(rc, "mp_layernorm")
    _count()


def embed_joints(x2d, w, b, spos, ln_g, ln_b, ln_eps, x_out, h_out, n_tokens, n_joints, c, dtype):
    rc = L.load().mp_embed_joints(L.ptr(x2d), L.ptr(w), L.ptr(b), L.ptr(spos), L.ptr(ln_g), L.ptr(ln_b), ln_eps, L.ptr(x_out),
                                  L.ptr(h_out), n_tokens, n_joints, c, dtype, L.stream_ptr())
    L.check(rc, "mp_embed_joints")
    _count()


def embed_segments(x2d, w, b, spos, ln_g, ln_b, ln_eps, x_out, h_out, n_frames, in_features, n_segments, c, dtype):
    rc = L.load().mp_embed_segments(L.ptr(x2d), L.ptr(w), L.ptr(b), L.ptr(spos), L.ptr(ln_g), L.ptr(ln_b), ln_eps, L.ptr(x_out),
                                    L.ptr(h_out), n_frames, in_features, n_segments, c, dtype, L.stream_ptr())
    L.check(rc, "mp_embed_segments")
    _count()


def attention(qkv, out, n_clips, n_frames, n_tok, c, n_heads, mode):
    rc = L.load().mp_attention(L.ptr(qkv), L.ptr(out), n_clips, n_frames, n_tok, c, n_heads, mode, DTYPE_CODE[qkv.dtype], L.stream_ptr())
    L.check(rc, "mp_attention")
    _count()
    return out


def heads_fwd(x, post_g, post_b, post_eps, hg, hb, hw, hbias, score_w, score_b, rot, logits, n_clips, n_frames, n_hyp, out_dim,
              with_score):
    rc = L.load().mp_heads_fwd(L.ptr(x), L.ptr(post_g), L.ptr(post_b), post_eps, L.ptr(hg), L.ptr(hb), L.ptr(hw), L.ptr(hbias),
                               L.ptr(score_w), L.ptr(score_b), L.ptr(rot), L.ptr(logits), n_clips, n_frames, n_hyp, out_dim,
                               int(with_score), L.stream_ptr())
    L.check(rc, "mp_heads_fwd")
    _count()


def heads_fwd16(xhat16, wf16, bf, score_w, score_b, rot, logits, workspace, n_clips, n_frames, n_hyp, out_dim, with_score):
    """K heads on the tensor cores: xhat16 [tokens, 512] (normalised, 16-bit) x folded weights wf16 [n_pad, 512] -> rot / logits."""
    rc = L.load().mp_heads_fwd16(L.ptr(xhat16), L.ptr(wf16), L.ptr(bf), L.ptr(score_w), L.ptr(score_b), L.ptr(rot), L.ptr(logits),
                                 L.ptr(workspace), workspace.numel() * workspace.element_size(), n_clips, n_frames, n_hyp, out_dim,
                                 int(with_score), wf16.shape[0], DTYPE_CODE[xhat16.dtype], L.stream_ptr())
    L.check(rc, "mp_heads_fwd16")
    _count(2)


def bones_head(x, post_g, post_b, post_eps, hg, hb, hw, hbias, bone_len, n_clips, n_frames, n_segments, c, workspace):
    rc = L.load().mp_bones_head(L.ptr(x), L.ptr(post_g), L.ptr(post_b), post_eps, L.ptr(hg), L.ptr(hb), L.ptr(hw), L.ptr(hbias),
                                L.ptr(bone_len), n_clips, n_frames, n_segments, c, L.ptr(workspace),
                                workspace.numel() * workspace.element_size(), L.stream_ptr())
    L.check(rc, "mp_bones_head")
    _count(2)
