"""Backward-pass wrappers over the C ABI (include/manipose_sm100.h, "backward (training) entry points").

The reference differentiates the MixSTE blocks with torch autograd (hpe/mh_so3_hpe/architectures/mix_ste.py:194-368); here
the backward of the trunk is an explicit reverse sweep (architectures/mix_ste.py::MixSTE._train_backward) over these
kernels, and the two autograd Functions below cover the small pieces around it (hypothesis heads, bone-length head).
No CPU fallback: every function needs CUDA tensors and the built library.
"""
from typing import Optional

import torch

from . import _lib as L
from . import ops

_zeros_cache = {}


def zeros_f32(n: int, device) -> torch.Tensor:
    """A shared read-only fp32 zero vector (bias operand of the backward GEMMs)."""
    key = (str(device), n)
    z = _zeros_cache.get(key)
    if z is None:
        z = torch.zeros(n, dtype=torch.float32, device=device)
        _zeros_cache[key] = z
    return z


def pad64(m: int) -> int:
    return (m + 63) // 64 * 64


def layernorm_bwd(x, gamma, eps, dy, dres, dx, dgamma, dbeta, dtype, dx16=None, rowscale=None, dx16_colsum=None):
    """dx = LN'(x)(dy) [+ dres]; dgamma / dbeta accumulate (None to skip).  dy fp32 or 16-bit.  dx16: also 16-bit(rowscale * dx), whose
    column sums are added to dx16_colsum (the bias gradient of the Linear that dx16 is the output gradient of)."""
    n_tokens, c = x.shape
    rc = L.load().mp_layernorm_bwd(L.ptr(x), L.ptr(gamma), float(eps), L.ptr(dy), int(dy.dtype != torch.float32), L.ptr(dres), L.ptr(dx),
                                   L.ptr(dgamma), L.ptr(dbeta), L.ptr(dx16), L.ptr(rowscale), L.ptr(dx16_colsum), n_tokens, c, dtype,
                                   L.stream_ptr())
    L.check(rc, "mp_layernorm_bwd")
    ops._count()
    return dx


def gelu_fwd(u, a):
    L.check(L.load().mp_gelu_fwd(L.ptr(u), L.ptr(a), u.numel(), ops.DTYPE_CODE[u.dtype], L.stream_ptr()), "mp_gelu_fwd")
    ops._count()
    return a


def gelu_bwd(u, da, du, colsum=None):
    """du = da * gelu'(u); colsum [C] += column sums of du (the fc1 bias gradient) in the same pass."""
    if colsum is None:
        L.check(L.load().mp_gelu_bwd(L.ptr(u), L.ptr(da), L.ptr(du), u.numel(), ops.DTYPE_CODE[u.dtype], L.stream_ptr()), "mp_gelu_bwd")
    else:
        m, c = u.shape
        L.check(L.load().mp_gelu_bwd_colsum(L.ptr(u), L.ptr(da), L.ptr(du), L.ptr(colsum), m, c, ops.DTYPE_CODE[u.dtype], L.stream_ptr()),
                "mp_gelu_bwd_colsum")
    ops._count()
    return du


def attention_bwd(qkv, o, dout, dqkv, n_clips, n_frames, n_tok, c, n_heads, mode, colsum=None):
    """dqkv from dout; colsum [3C] += column sums of dqkv (the qkv bias gradient)."""
    rc = L.load().mp_attention_bwd(L.ptr(qkv), L.ptr(o), L.ptr(dout), L.ptr(dqkv), L.ptr(colsum), n_clips, n_frames, n_tok, c, n_heads, mode,
                                   ops.DTYPE_CODE[qkv.dtype], L.stream_ptr())
    L.check(rc, "mp_attention_bwd")
    ops._count()
    return dqkv


def transpose16(src, dst, colsum=None):
    """dst[C, Mpad] = src[M, C]^T zero padded; colsum[C] += column sums of src."""
    m, c = src.shape
    rc = L.load().mp_transpose16(L.ptr(src), L.ptr(dst), L.ptr(colsum), m, c, dst.shape[1], ops.DTYPE_CODE[src.dtype], L.stream_ptr())
    L.check(rc, "mp_transpose16")
    ops._count()
    return dst


def group_rowsum(x, out, div, mod):
    n_rows, c = x.shape
    L.check(L.load().mp_group_rowsum(L.ptr(x), L.ptr(out), n_rows, c, div, mod, L.stream_ptr()), "mp_group_rowsum")
    ops._count()


def small_wgrad(dy, inp, dw, db):
    n_rows, n_out = dy.shape
    L.check(L.load().mp_small_wgrad(L.ptr(dy), L.ptr(inp), L.ptr(dw), L.ptr(db), n_rows, n_out, inp.shape[1], L.stream_ptr()), "mp_small_wgrad")
    ops._count()


def residual_rowscale(x, y16, s, out):
    n_tokens, c = x.shape
    rc = L.load().mp_residual_rowscale(L.ptr(x), L.ptr(y16), L.ptr(s), L.ptr(out), n_tokens, c, ops.DTYPE_CODE[y16.dtype], L.stream_ptr())
    L.check(rc, "mp_residual_rowscale")
    ops._count()
    return out


def cast_rowscale(g, s, out16):
    n_tokens, c = g.shape
    rc = L.load().mp_cast_rowscale(L.ptr(g), L.ptr(s), L.ptr(out16), n_tokens, c, ops.DTYPE_CODE[out16.dtype], L.stream_ptr())
    L.check(rc, "mp_cast_rowscale")
    ops._count()
    return out16


def adam_step(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0, step_dev=None, lr_dev=None):
    """``step`` counts from 1 on the host; with ``step_dev`` (int64 device scalar) the count lives on the device instead, and with
    ``lr_dev`` (fp32 device scalar) so does the learning rate."""
    rc = L.load().mp_adam_step(L.ptr(param), L.ptr(grad), L.ptr(exp_avg), L.ptr(exp_avg_sq), param.numel(), float(lr), float(beta1),
                               float(beta2), float(eps), float(weight_decay), int(step), L.ptr(step_dev), L.ptr(lr_dev), float(grad_scale),
                               L.stream_ptr())
    L.check(rc, "mp_adam_step")
    ops._count(1 if step_dev is None else 2)


def grad_of(p: torch.Tensor) -> torch.Tensor:
    """The fp32 buffer the backward kernels accumulate into (``p.grad``; created zeroed when missing, like autograd would)."""
    if p.grad is None:
        p.grad = torch.zeros_like(p, memory_format=torch.contiguous_format)
    return p.grad


def colsum16(src16, colsum):
    m, c = src16.shape
    L.check(L.load().mp_colsum16(L.ptr(src16), L.ptr(colsum), m, c, ops.DTYPE_CODE[src16.dtype], L.stream_ptr()), "mp_colsum16")
    ops._count()


def wgrad(dy16, act16, dw, db):
    """dw[N,K] += dy16[M,N]^T act16[M,K]; db[N] += column sums of dy16 (db None to skip).

    tcgen05 path: both operands are read in place as MN-major UMMA operands (rows = tokens = the contraction index), the token
    contraction is split over the SMs and the partial tiles are added into the fp32 gradient with TMA reduce stores."""
    m, n = dy16.shape
    k = act16.shape[1]
    if db is not None:
        colsum16(dy16, db)
    rc = L.load().mp_wgrad(L.ptr(dy16), L.ptr(act16), L.ptr(dw), m, n, k, ops.DTYPE_CODE[dy16.dtype], L.stream_ptr())
    L.check(rc, "mp_wgrad")
    ops._count()


def dgrad(dy16, w_t16, out16):
    """out16[M,K] = dy16[M,N] @ W[N,K], with the transposed 16-bit weight shadow w_t16 [K,N]."""
    return ops.linear(dy16, w_t16, zeros_f32(max(w_t16.shape[0], 2048), dy16.device), out16, L.MP_EPI_BIAS)


# ------------------------------------------------------------------------------------------------ small autograd pieces
class LayerNormFn(torch.autograd.Function):
    """y = LayerNorm(x fp32 [M,C]; gamma, beta, eps) as fp32 (``out16`` None) or as a 16-bit GEMM operand."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps, out16):
        x = ops._f32(x)
        m, c = x.shape
        if out16 is None:
            y = torch.empty_like(x)
            ops.layernorm(x, y, None, post=(gamma, beta), post_eps=eps)
        else:
            y = torch.empty((m, c), dtype=ops.TORCH_DTYPE[out16], device=x.device)
            ops.layernorm(x, None, y, ln=(gamma, beta), ln_eps=eps, dtype=out16)
        ctx.save_for_backward(x, gamma)
        ctx.eps, ctx.code = eps, out16 if out16 is not None else L.MP_DTYPE_BF16
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        need_affine = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        dg = torch.zeros_like(gamma) if need_affine else None
        db = torch.zeros_like(gamma) if need_affine else None
        layernorm_bwd(x, gamma, ctx.eps, dy, None, dx, dg, db, ctx.code)
        return dx, dg, db, None, None


class LinearF32Fn(torch.autograd.Function):
    """y fp32 [M,N] = a16 [M,K] @ W[N,K]^T + b  with W, b fp32 parameters (cast to 16-bit inside); N % 128 == 0.
    Used for the (zero-padded) hypothesis heads and the bone-length head, whose outputs must stay fp32."""

    @staticmethod
    def forward(ctx, a16, w, b):
        code = ops.DTYPE_CODE[a16.dtype]
        w16 = ops.cast16(w.detach(), code)
        m, n = a16.shape[0], w.shape[0]
        y = torch.zeros((m, n), dtype=torch.float32, device=a16.device)
        ops.linear(a16, w16, ops._f32(b.detach()), y, L.MP_EPI_RESIDUAL, resid=y)
        ctx.save_for_backward(a16, w16)
        return y

    @staticmethod
    def backward(ctx, dy):
        a16, w16 = ctx.saved_tensors
        code = ops.DTYPE_CODE[a16.dtype]
        m, k = a16.shape
        n = w16.shape[0]
        dy16 = ops.cast16(dy.contiguous(), code)
        dw = torch.zeros((n, k), dtype=torch.float32, device=a16.device)
        db = torch.zeros(n, dtype=torch.float32, device=a16.device)
        wgrad(dy16, a16, dw, db)
        w_t = torch.empty((k, n), dtype=a16.dtype, device=a16.device)
        transpose16(w16, w_t)
        da = torch.empty((m, k), dtype=a16.dtype, device=a16.device)
        dgrad(dy16, w_t, da)
        return da, dw, db


def stacked_grads(params) -> Optional[torch.Tensor]:
    """One strided view [K, *shape] over the gradients of K same-shaped parameters when they sit at a constant stride in one
    storage (the flat gradient buffer of ``optim.FlatParameters``: the K hypothesis heads are laid out one after the other), so K
    per-parameter accumulations become one launch.  None when the layout does not allow it."""
    grads = [grad_of(p) for p in params]
    g0 = grads[0]
    if len(grads) == 1:
        return g0.unsqueeze(0)
    base = g0.untyped_storage().data_ptr()
    if any(g.untyped_storage().data_ptr() != base or not g.is_contiguous() or g.shape != g0.shape for g in grads):
        return None
    step = grads[1].storage_offset() - g0.storage_offset()
    if step <= 0 or any(g.storage_offset() != g0.storage_offset() + i * step for i, g in enumerate(grads)):
        return None
    return torch.as_strided(g0, (len(grads),) + tuple(g0.shape), (step,) + tuple(g0.stride()), g0.storage_offset())


def accumulate_stacked(params, stacked: torch.Tensor) -> None:
    """grad(params[k]) += stacked[k] — one launch when the gradients form a strided stack, K otherwise."""
    view = stacked_grads(params)
    if view is not None:
        view.add_(stacked.reshape(view.shape))
    else:
        for k, p in enumerate(params):
            grad_of(p).add_(stacked[k].reshape(p.shape))


def _head_tables(heads):
    """Device tables int64 [6][K] of the heads' parameter / gradient pointers (mp_heads_fold / mp_heads_bwd_pack / mp_heads_unfold),
    rebuilt only when a tensor moved (the optimizer keeps parameters and gradients in flat buffers, so they do not)."""
    kinds = (lambda h: h.norm.weight, lambda h: h.norm.bias, lambda h: h.prediction_head.weight, lambda h: h.prediction_head.bias,
             lambda h: h.score_head.weight, lambda h: h.score_head.bias)
    params = [f(h) for f in kinds for h in heads]
    key = tuple(p.data_ptr() for p in params) + tuple(grad_of(p).data_ptr() for p in params)
    cached = getattr(heads[0], "_mp_tables", None)
    if cached is None or cached[0] != key:
        dev = params[0].device
        n = len(params)
        cached = (key, torch.tensor(key[:n], dtype=torch.int64, device=dev), torch.tensor(key[n:], dtype=torch.int64, device=dev))
        heads[0]._mp_tables = cached
    return cached[1], cached[2]


class FoldedHeadsFn(torch.autograd.Function):
    """The K hypothesis heads of RMCLRotMixSTE (rmcl_manifold_mix_ste.py:251-298) over a shared normalised input yhat [M, C] (16-bit):
    LN_k(y) = yhat * gamma_k + beta_k, so all heads are ONE Linear with folded parameters W_k * gamma_k, W_k beta_k + b_k (zero-padded to
    128 outputs for the tensor-core kernel, fp32 output), followed by the J-term score dot product.  Parameters are read from the
    module, not passed through autograd; folding, the packing of the output gradient and the unfolding of the parameter gradients are one
    kernel each (csrc/heads_train.cu) around the GEMMs, and the gradients are accumulated straight into ``p.grad`` like the trunk's
    reverse sweep does: 3 launches forward, 5 backward (the torch version of this glue was ~55)."""

    @staticmethod
    def forward(ctx, yhat16, heads, b, l, j, out_dim, anchor):   # anchor: a head parameter, only there so that the outputs require grad
        k, d1, c = len(heads), out_dim + 1, yhat16.shape[1]
        code = ops.DTYPE_CODE[yhat16.dtype]
        dev = yhat16.device
        params, grads = _head_tables(heads)
        n_pad = (k * d1 + 127) // 128 * 128
        m = yhat16.shape[0]
        wf16 = torch.empty((n_pad, c), dtype=yhat16.dtype, device=dev)
        wt16 = torch.empty((c, n_pad), dtype=yhat16.dtype, device=dev)
        bf = torch.empty(n_pad, dtype=torch.float32, device=dev)
        sw = torch.empty((k, j), dtype=torch.float32, device=dev)
        sb = torch.empty(k, dtype=torch.float32, device=dev)
        L.check(L.load().mp_heads_fold(L.ptr(params), k, out_dim, c, n_pad, L.ptr(wf16), L.ptr(wt16), L.ptr(bf), L.ptr(sw), L.ptr(sb), code,
                                       L.stream_ptr()), "mp_heads_fold")
        ops._count()
        y = torch.empty((m, n_pad), dtype=torch.float32, device=dev)     # the GEMM output: the score embeddings are needed again
        rot = torch.empty((b, k, l, j, out_dim), dtype=torch.float32, device=dev)
        logits = torch.empty((b, k, l), dtype=torch.float32, device=dev)
        ops.heads_fwd16(yhat16, wf16, bf, sw, sb, rot, logits, y, b, l, k, out_dim, True)
        ctx.tables, ctx.dims = (params, grads), (b, l, j, k, d1, c, n_pad, out_dim)
        ctx.save_for_backward(yhat16, wt16, y, sw)
        return rot, logits

    @staticmethod
    def backward(ctx, d_rot, d_logits):
        yhat16, wt16, y, sw = ctx.saved_tensors
        params, grads = ctx.tables
        b, l, j, k, d1, c, n_pad, out_dim = ctx.dims
        code = ops.DTYPE_CODE[yhat16.dtype]
        dev = yhat16.device
        m = yhat16.shape[0]
        if d_rot is None:
            d_rot = torch.zeros((b, k, l, j, out_dim), dtype=torch.float32, device=dev)
        if d_logits is None:
            d_logits = torch.zeros((b, k, l), dtype=torch.float32, device=dev)
        buf = torch.zeros(n_pad * c + n_pad, dtype=torch.float32, device=dev)
        dwf, dbf = buf[:n_pad * c].view(n_pad, c), buf[n_pad * c:]
        dy16 = torch.empty((m, n_pad), dtype=yhat16.dtype, device=dev)
        rc = L.load().mp_heads_bwd_pack(L.ptr(ops._f32(d_rot)), L.ptr(ops._f32(d_logits)), L.ptr(y), L.ptr(sw), L.ptr(dy16), L.ptr(dbf), L.ptr(grads),
                                        b, l, k, out_dim, n_pad, code, L.stream_ptr())
        L.check(rc, "mp_heads_bwd_pack")
        ops._count()
        wgrad(dy16, yhat16, dwf, None)
        da = torch.empty((m, c), dtype=yhat16.dtype, device=dev)
        dgrad(dy16, wt16, da)
        L.check(L.load().mp_heads_unfold(L.ptr(params), L.ptr(grads), L.ptr(dwf), L.ptr(dbf), k, out_dim, c, L.stream_ptr()), "mp_heads_unfold")
        ops._count()
        return da, None, None, None, None, None, None


def layer_norm(x, gamma, beta, eps, out16: Optional[int] = None):
    return LayerNormFn.apply(x, gamma, beta, eps, out16)


def linear_f32(a16, w, b):
    return LinearF32Fn.apply(a16, w, b)
