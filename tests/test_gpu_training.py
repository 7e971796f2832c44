"""Backward (training) path: each kernel against torch autograd of the same op in fp32, then the gradients of the whole model
+ default objective against autograd through the fp32 CPU oracle ("autograd goldens for every backward kernel
are obtainable"), then Adam against torch.optim.Adam.

Tolerances: activation gradients travel as 16-bit tensor-core operands (as under bf16/fp16 autocast), so per-parameter
gradients are compared by relative L2 norm: 16-bit rounding noise averages out over the token reduction of a weight
gradient; the limits below are ~1.5-2x the measured values (printed by the test)."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import manipose_oracle as O

pytestmark = pytest.mark.gpu

DT = {"bf16": torch.bfloat16, "fp16": torch.float16}
CODE = {"bf16": 0, "fp16": 1}
RTOL = {"bf16": 1e-2, "fp16": 2e-3}


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


@pytest.mark.parametrize("c", [512, 128])
@pytest.mark.parametrize("dy16", [False, True])
@pytest.mark.parametrize("with_res", [False, True])
def test_layernorm_bwd_vs_autograd(c, dy16, with_res):
    from manipose_b200 import train_ops as T, _lib as L
    gen = torch.Generator(device="cuda").manual_seed(c + dy16)
    m = 1037
    x = (torch.randn(m, c, generator=gen, device="cuda") * 1.7 + 0.3).requires_grad_()
    gamma = (1 + 0.1 * torch.randn(c, generator=gen, device="cuda")).requires_grad_()
    beta = (0.1 * torch.randn(c, generator=gen, device="cuda")).requires_grad_()
    dy = torch.randn(m, c, generator=gen, device="cuda")
    if dy16:
        dy = dy.bfloat16()
    res = torch.randn(m, c, generator=gen, device="cuda") if with_res else None
    F.layer_norm(x, (c,), gamma, beta, 1e-6).backward(dy.float())
    dx = torch.full((m, c), float("nan"), device="cuda")
    dg, db = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
    T.layernorm_bwd(x.detach(), gamma.detach(), 1e-6, dy, res, dx, dg, db, L.MP_DTYPE_BF16)
    want = x.grad + (res if with_res else 0)
    torch.testing.assert_close(dx, want, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(dg, gamma.grad, rtol=1e-4, atol=2e-3)
    torch.testing.assert_close(db, beta.grad, rtol=1e-4, atol=2e-3)
    # in place over the residual gradient (how the sweep uses it) and without affine
    if with_res:
        buf = res.clone()
        T.layernorm_bwd(x.detach(), gamma.detach(), 1e-6, dy, buf, buf, None, None, L.MP_DTYPE_BF16)
        torch.testing.assert_close(buf, want, rtol=1e-4, atol=1e-4)
    # the 16-bit, row-scaled copy for the next GEMM and its column sums (that Linear's bias gradient), taken in the same pass
    scale = 0.5 + torch.rand(m, generator=gen, device="cuda")
    dx16 = torch.empty((m, c), dtype=torch.bfloat16, device="cuda")
    col = torch.full((c,), 3.0, device="cuda")
    T.layernorm_bwd(x.detach(), gamma.detach(), 1e-6, dy, res, dx, None, None, L.MP_DTYPE_BF16, dx16=dx16, rowscale=scale, dx16_colsum=col)
    torch.testing.assert_close(dx16.float(), want * scale[:, None], rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(col - 3.0, (want * scale[:, None]).sum(0), rtol=1e-3, atol=2e-2)


@pytest.mark.parametrize("dtype", ["bf16", "fp16"])
def test_gelu_fwd_bwd(dtype):
    from manipose_b200 import train_ops as T
    td = DT[dtype]
    gen = torch.Generator(device="cuda").manual_seed(1)
    u = (torch.randn(3000, 1024, generator=gen, device="cuda") * 2).to(td)
    da = torch.randn(3000, 1024, generator=gen, device="cuda").to(td)
    uf = u.float().requires_grad_()
    ref = F.gelu(uf)
    ref.backward(da.float())
    a, du = torch.empty_like(u), torch.empty_like(u)
    T.gelu_fwd(u, a)
    T.gelu_bwd(u, da, du)
    torch.testing.assert_close(a.float(), ref.detach(), rtol=RTOL[dtype], atol=RTOL[dtype])
    torch.testing.assert_close(du.float(), uf.grad, rtol=RTOL[dtype], atol=RTOL[dtype])
    # the same with the column sums of du (the fc1 bias gradient) taken in the same pass; ragged row count, narrow matrix
    for rows, cols in ((3000, 1024), (777, 256), (5, 8)):
        du2, col = torch.empty((rows, cols), dtype=td, device="cuda"), torch.full((cols,), -2.0, device="cuda")
        T.gelu_bwd(u[:rows, :cols].contiguous(), da[:rows, :cols].contiguous(), du2, colsum=col)
        assert torch.equal(du2, du[:rows, :cols])
        torch.testing.assert_close(col + 2.0, du2.float().sum(0), rtol=1e-4, atol=2e-2)


def _attn(qkv, n_clips, n_frames, n_tok, c, heads, temporal):
    hd = c // heads
    x = qkv.reshape(n_clips, n_frames, n_tok, 3, heads, hd)
    q, k, v = x[:, :, :, 0], x[:, :, :, 1], x[:, :, :, 2]
    if temporal:
        q, k, v = (t.permute(0, 2, 3, 1, 4) for t in (q, k, v))
    else:
        q, k, v = (t.permute(0, 1, 3, 2, 4) for t in (q, k, v))
    o = ((q @ k.transpose(-1, -2)) * hd ** -0.5).softmax(-1) @ v
    o = o.permute(0, 3, 1, 2, 4) if temporal else o.permute(0, 1, 3, 2, 4)
    return o.reshape(n_clips * n_frames * n_tok, c)


@pytest.mark.parametrize("dtype", ["bf16", "fp16"])
@pytest.mark.parametrize("n_clips,n_frames,n_tok,c,temporal", [
    (3, 27, 17, 512, True), (3, 27, 17, 512, False), (1, 243, 17, 512, True), (2, 243, 17, 512, False), (2, 81, 16, 128, True),
    (2, 27, 16, 128, False), (1, 1, 17, 512, True), (5, 9, 17, 512, False), (2, 81, 17, 512, True), (1, 128, 17, 512, True),
    (2, 50, 17, 512, True), (1, 3, 17, 512, False)])
def test_attention_bwd_vs_autograd(n_clips, n_frames, n_tok, c, temporal, dtype):
    from manipose_b200 import ops, train_ops as T
    td = DT[dtype]
    gen = torch.Generator(device="cuda").manual_seed(n_frames + c)
    n = n_clips * n_frames * n_tok
    qkv = (torch.randn(n, 3 * c, generator=gen, device="cuda") * 1.2).to(td)
    do = torch.randn(n, c, generator=gen, device="cuda").to(td)
    o = torch.empty((n, c), dtype=td, device="cuda")
    ops.attention(qkv, o, n_clips, n_frames, n_tok, c, 8, 1 if temporal else 0)
    qf = qkv.float().requires_grad_()
    _attn(qf, n_clips, n_frames, n_tok, c, 8, temporal).backward(do.float())
    dqkv = torch.full((n, 3 * c), float("nan"), dtype=td, device="cuda")
    col = torch.full((3 * c,), 1.5, device="cuda")
    T.attention_bwd(qkv, o, do, dqkv, n_clips, n_frames, n_tok, c, 8, 1 if temporal else 0, colsum=col)
    torch.cuda.synchronize()
    assert not torch.isnan(dqkv.float()).any()
    want_col = dqkv.float().sum(0)                 # the qkv bias gradient, taken by the same kernel
    torch.testing.assert_close(col - 1.5, want_col, rtol=2e-3, atol=2e-3 * float(want_col.abs().max()) + 1e-3)
    floor = 0.05 * float(qf.grad.norm())     # L = 1: dq = dk = 0 exactly, ours carries the 16-bit rounding of O
    for name, sl in (("dq", slice(0, c)), ("dk", slice(c, 2 * c)), ("dv", slice(2 * c, 3 * c))):
        err = float((dqkv[:, sl].float() - qf.grad[:, sl]).norm())
        assert err <= 2 * RTOL[dtype] * max(float(qf.grad[:, sl].norm()), floor), name


@pytest.mark.parametrize("dtype", ["bf16", "fp16"])
@pytest.mark.parametrize("m,n,k", [(13770, 1536, 512), (13770, 512, 1024), (1000, 384, 128), (77, 128, 256), (12960, 128, 128)])
def test_wgrad_dgrad_vs_fp32(m, n, k, dtype):
    """dW += dY^T X (MN-major UMMA operands, split-K with TMA reduce stores, in-place fp32 accumulation), db, and dX = dY W."""
    from manipose_b200 import train_ops as T
    td = DT[dtype]
    gen = torch.Generator(device="cuda").manual_seed(m + n)
    dy = torch.randn(m, n, generator=gen, device="cuda").to(td)
    x = torch.randn(m, k, generator=gen, device="cuda").to(td)
    w = (torch.randn(n, k, generator=gen, device="cuda") / math.sqrt(k)).to(td)
    dw0 = torch.randn(n, k, generator=gen, device="cuda")
    db0 = torch.randn(n, generator=gen, device="cuda")
    dw, db = dw0.clone(), db0.clone()
    T.wgrad(dy, x, dw, db)
    torch.testing.assert_close(dw, dw0 + dy.float().t() @ x.float(), rtol=2e-3, atol=2e-2)
    torch.testing.assert_close(db, db0 + dy.float().sum(0), rtol=2e-3, atol=2e-2)
    w_t = T.transpose16(w, torch.empty((k, n), dtype=td, device="cuda"))
    assert torch.equal(w_t, w.t().contiguous())
    dx = torch.empty((m, k), dtype=td, device="cuda")
    T.dgrad(dy, w_t, dx)
    torch.testing.assert_close(dx.float(), dy.float() @ w.float(), rtol=RTOL[dtype], atol=RTOL[dtype] * math.sqrt(n) / 8)


def test_small_reductions():
    from manipose_b200 import train_ops as T
    gen = torch.Generator(device="cuda").manual_seed(5)
    b, t, j, c = 3, 27, 17, 512
    dx = torch.randn(b * t * j, c, generator=gen, device="cuda")
    spos, tpos = torch.zeros(j, c, device="cuda"), torch.zeros(t, c, device="cuda")
    T.group_rowsum(dx, spos, 1, j)
    T.group_rowsum(dx, tpos, j, t)
    v = dx.view(b, t, j, c)
    torch.testing.assert_close(spos, v.sum((0, 1)), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(tpos, v.sum((0, 2)), rtol=1e-4, atol=1e-4)
    inp = torch.randn(b * t * j, 2, generator=gen, device="cuda")
    dw, db = torch.zeros(c, 2, device="cuda"), torch.zeros(c, device="cuda")
    T.small_wgrad(dx, inp, dw, db)
    torch.testing.assert_close(dw, dx.t() @ inp, rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(db, dx.sum(0), rtol=1e-4, atol=1e-3)
    dy = torch.randn(b * t, 2048, generator=gen, device="cuda")
    inp = torch.randn(b * t, 34, generator=gen, device="cuda")
    dw, db = torch.zeros(2048, 34, device="cuda"), torch.zeros(2048, device="cuda")
    T.small_wgrad(dy, inp, dw, db)
    torch.testing.assert_close(dw, dy.t() @ inp, rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(db, dy.sum(0), rtol=1e-4, atol=1e-3)
    # stochastic-depth glue
    x = torch.randn(1000, 512, generator=gen, device="cuda")
    y = torch.randn(1000, 512, generator=gen, device="cuda").bfloat16()
    s = (torch.rand(1000, generator=gen, device="cuda") > 0.3).float() / 0.7
    out = T.residual_rowscale(x, y, s, torch.empty_like(x))
    torch.testing.assert_close(out, x + s[:, None] * y.float(), rtol=1e-6, atol=1e-6)
    g16 = T.cast_rowscale(x, s, torch.empty_like(y))
    assert torch.equal(g16, (x * s[:, None]).bfloat16())
    assert torch.equal(T.cast_rowscale(x, None, torch.empty_like(y)), x.bfloat16())


def test_adam_vs_torch():
    from manipose_b200 import train_ops as T
    gen = torch.Generator(device="cuda").manual_seed(9)
    p = torch.randn(100003, generator=gen, device="cuda")
    ref = p.clone().requires_grad_()
    opt = torch.optim.Adam([ref], lr=4e-5, weight_decay=1e-6)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 6):
        g = torch.randn(100003, generator=gen, device="cuda")
        ref.grad = g.clone()
        opt.step()
        T.adam_step(p, g * 2.0, m, v, 4e-5, 0.9, 0.999, 1e-8, 1e-6, step, grad_scale=0.5)
    torch.testing.assert_close(p, ref.detach(), rtol=1e-6, atol=1e-7)


def _model_from_sd(sd, num_frame, n_hyp, dtype, **kw):
    import manipose_b200 as mb
    m = mb.RMCLManifoldMixSTE(mb.h36m17_skeleton(), num_frame=num_frame, n_hyp=n_hyp, **kw)
    m.load_state_dict(sd)
    return m.cuda().set_compute_dtype(dtype)


# relative L2 of a parameter gradient vs fp32 autograd through the oracle: (median over the 290 tensors, worst tensor).
# Measured on B200: fp16 median 3.5e-3..4.7e-3 / worst 8e-3; bf16 median 2.0e-2..4.6e-2 / worst 0.19 (a score head).
#
# What the bf16 figure is: the rounding of ~100 chained 16-bit activations (8 significand bits), not the backward kernels and not
# winner flips.  Two controls are part of the test:
#   * the SAME reference model under ``torch.autocast(bfloat16)`` on the CPU (PyTorch's own bf16 kernels, forward and autograd) is
#     compared with the fp32 oracle in the same way: it measures 5.7e-2 median / 0.28 worst at T=27 (fp16: 4.2e-3 / 1.2e-2) — the
#     sm_100a kernels must be NO WORSE than that (median and 90th percentile within 1.25 x), at both dtypes;
#   * ``test_model_gradients_without_winner_flips`` removes the flips by construction and shows the error does not go down.
GRAD_TOL = {"bf16": (8e-2, 4e-1), "fp16": (1e-2, 3e-2)}


def _autocast_oracle_grads(x, y, sd, dtype):
    """Parameter gradients of the reference model under torch.autocast(dtype) on the CPU (PyTorch's own 16-bit kernels + autograd)."""
    sd_ac = {k: v.clone().requires_grad_() for k, v in sd.items()}
    with torch.autocast("cpu", dtype=torch.bfloat16 if dtype == "bf16" else torch.float16):
        poses, scores = O.rmcl_forward(x, sd_ac)
    loss, _ = O.training_loss(poses.float(), scores.float(), y)
    loss.backward()
    return {k: v.grad for k, v in sd_ac.items()}


def _grad_report(tag, model, sd_ref, sd_autocast):
    rels = {name: _rel(p.grad.cpu(), sd_ref[name].grad) for name, p in model.named_parameters()}
    vals = sorted(rels.values())
    ac = sorted(_rel(sd_autocast[name], sd_ref[name].grad) for name in rels)
    q = lambda v, f: v[int(f * (len(v) - 1))]
    print(f"[{tag}] grad rel-L2 vs fp32 oracle: median {q(vals, .5):.2e}, p90 {q(vals, .9):.2e}, max {vals[-1]:.2e} ({max(rels, key=rels.get)}); "
          f"torch autocast on the CPU: median {q(ac, .5):.2e}, p90 {q(ac, .9):.2e}, max {ac[-1]:.2e}")
    return rels, vals, ac, q


@pytest.mark.parametrize("dtype", ["fp16", "bf16"])
@pytest.mark.parametrize("T_,K,B", [(27, 5, 2), (9, 2, 3)])
def test_model_gradients_vs_oracle_autograd(T_, K, B, dtype):
    """BASELINE config 4 shape (T=27, K=5, default widths and depth) at a small batch: loss value and all 290 parameter gradients
    of forward + default objective (wta + 0.1 bce + 2 velocity + 0.5 smoothness) vs torch autograd through the fp32 CPU oracle,
    and against PyTorch's own 16-bit autocast of the same model as the yardstick of what the format allows."""
    from manipose_b200 import metrics
    sd = O.make_state_dict(num_frame=T_, n_hyp=K, seed=11)
    gen = torch.Generator().manual_seed(21)
    x = 0.3 * torch.randn(B, T_, 17, 2, generator=gen)
    y = 0.3 * torch.randn(B, T_, 17, 3, generator=gen)
    y[:, :, 0] = 0
    sd_ref = {k: v.clone().requires_grad_() for k, v in sd.items()}
    poses_ref, scores_ref = O.rmcl_forward(x, sd_ref)
    loss_ref, _ = O.training_loss(poses_ref, scores_ref, y)
    loss_ref.backward()
    sd_ac = _autocast_oracle_grads(x, y, sd, dtype)

    m = _model_from_sd(sd, T_, K, dtype, drop_path_rate=0.0).train()
    poses, scores = m(x.cuda())
    assert poses.requires_grad and scores.requires_grad
    loss, terms = metrics.losses.training_loss(poses, scores, y.cuda())
    loss.backward()
    torch.cuda.synchronize()
    assert abs(float(loss.detach()) - float(loss_ref.detach())) <= 2e-2 * abs(float(loss_ref.detach()))
    for name, p in m.named_parameters():
        assert p.grad is not None and sd_ref[name].grad is not None, name
    typical, worst = GRAD_TOL[dtype]
    rels, vals, ac, q = _grad_report(f"{dtype} T={T_}", m, sd_ref, sd_ac)
    bad = {k: v for k, v in rels.items() if not v <= worst}
    assert not bad, bad
    assert q(vals, .5) <= typical
    # no worse than PyTorch's own autocast of the reference model at this dtype
    assert q(vals, .5) <= 1.25 * q(ac, .5) and q(vals, .9) <= 1.25 * q(ac, .9)


# The same comparison with the winner-takes-all flips REMOVED by construction and with the weight rounding taken out:
#   * the target of every frame is one of the oracle's own hypotheses (k* = (clip + frame) mod K) with every joint displaced by 0.5
#     in a random direction: the winner stays a wide margin ahead of the other K - 1 hypotheses (smallest margin 0.14 against a
#     forward error of a few 1e-3) while the residuals stay O(0.5), so the unit-vector gradients of the L2 terms are as well
#     conditioned as with a random target (a target AT a hypothesis would put the objective at a minimum, where any forward error
#     dominates the gradient); the test first proves that no winner moved;
#   * the block GEMM weights are made representable in the 16-bit format before BOTH runs (the oracle linearises at the weights the
#     tensor cores multiply by), and the K heads get O(1) LayerNorm biases (the folded-head backward has a term in beta that is
#     zero at the default init).
# Measured on B200: fp16 median 6.3e-3 / worst 2.5e-2; bf16 median 9.9e-2 .. 1.6e-1 / worst 0.37 .. 0.59 depending on the build (any
# change of an fp32 summation order in the forward moves it) against 1.2e-1 / 0.25 for PyTorch's own bf16 autocast of the same model —
# NOT smaller than with flips: the bf16 figure is the rounding of ~100 chained 8-bit-significand activations (the O(1) head-norm
# biases of this construction amplify it), which is why the tight statement about the backward kernels is the fp16 one.
GRAD_TOL_NO_FLIPS = {"bf16": (2.5e-1, 8e-1), "fp16": (1.2e-2, 5e-2)}


@pytest.mark.parametrize("dtype", ["fp16", "bf16"])
def test_model_gradients_without_winner_flips(dtype):
    from manipose_b200 import metrics, ops
    T_, K, B = 27, 5, 2
    sd = O.make_state_dict(num_frame=T_, n_hyp=K, seed=11)
    td = torch.bfloat16 if dtype == "bf16" else torch.float16
    gen = torch.Generator().manual_seed(5)
    for name in list(sd):
        if ".head." in name and name.endswith("norm.bias"):
            sd[name] = torch.randn(sd[name].shape, generator=gen)
        if "blocks." in name and name.endswith(("qkv.weight", "proj.weight", "fc1.weight", "fc2.weight")):
            sd[name] = sd[name].to(td).float()
    x = 0.3 * torch.randn(B, T_, 17, 2, generator=gen)
    with torch.no_grad():
        poses0, _ = O.rmcl_forward(x, sd)
    pick = (torch.arange(B)[:, None] + torch.arange(T_)[None, :]) % K
    offset = torch.randn(B, T_, 17, 3, generator=gen)
    offset = 0.5 * offset / offset.norm(dim=-1, keepdim=True)
    y = poses0[torch.arange(B)[:, None], pick, torch.arange(T_)[None, :]] + offset
    y[:, :, 0] = 0
    sd_ref = {k: v.clone().requires_grad_() for k, v in sd.items()}
    poses_ref, scores_ref = O.rmcl_forward(x, sd_ref)
    loss_ref, _ = O.training_loss(poses_ref, scores_ref, y)
    loss_ref.backward()
    sd_ac = _autocast_oracle_grads(x, y, sd, dtype)

    m = _model_from_sd(sd, T_, K, dtype, drop_path_rate=0.0).train()
    poses, scores = m(x.cuda())
    _, idx = ops.wta_fwd(poses.detach(), y.cuda(), metrics.losses.STANDARD_H36M_WEIGHTS, False)
    assert torch.equal(idx.cpu(), pick), "a winner moved: the margin of this test is too small"
    loss, _ = metrics.losses.training_loss(poses, scores, y.cuda())
    loss.backward()
    torch.cuda.synchronize()
    assert abs(float(loss.detach()) - float(loss_ref.detach())) <= 2e-2 * abs(float(loss_ref.detach()))
    typical, worst = GRAD_TOL_NO_FLIPS[dtype]
    rels, vals, ac, q = _grad_report(f"{dtype} no flips", m, sd_ref, sd_ac)
    bad = {k: v for k, v in rels.items() if not v <= worst}
    assert not bad, bad
    assert q(vals, .5) <= typical
    # (the bone-length backbone's tensors, a tenth of the list, sit at 2.5e-2 at fp16 here against autocast's 1.6e-2: hence 2 x at p90)
    slack = 1.25 if dtype == "fp16" else 2.0
    assert q(vals, .5) <= slack * q(ac, .5) and q(vals, .9) <= 2.0 * q(ac, .9)


def test_eval_and_training_forward_agree():
    """The differentiable trunk (separate kernels, tape) and the fused inference trunk compute the same function."""
    sd = O.make_state_dict(num_frame=27, n_hyp=5, seed=4)
    x = 0.3 * torch.randn(3, 27, 17, 2, generator=torch.Generator().manual_seed(2)).cuda()
    m = _model_from_sd(sd, 27, 5, "fp16", drop_path_rate=0.0).eval()
    with torch.no_grad():
        p0, s0 = m(x)
    p1, s1 = m(x)          # grad enabled -> tape path
    assert p1.requires_grad
    assert float((p1 - p0).norm() / p0.norm()) <= 5e-3
    assert float((s1 - s0).abs().max()) <= 2e-3


def test_training_step_reduces_loss_and_droppath_runs():
    """A few FusedAdam steps on one batch with stochastic depth on (drop_path_rate 0.1, the drivers' value): finite gradients,
    the objective goes down, and the weight shadows follow the optimizer."""
    import manipose_b200 as mb
    from manipose_b200 import metrics
    from manipose_b200.optim import FusedAdam
    torch.manual_seed(0)
    m = mb.RMCLManifoldMixSTE(mb.h36m17_skeleton(), num_frame=27, n_hyp=5, drop_path_rate=0.1).cuda().train()
    opt = FusedAdam(m, lr=2e-4, weight_decay=1e-6)
    gen = torch.Generator().manual_seed(3)
    x = (0.3 * torch.randn(6, 27, 17, 2, generator=gen)).cuda()
    y = 0.3 * torch.randn(6, 27, 17, 3, generator=gen)
    y[:, :, 0] = 0
    y = y.cuda()
    losses = []
    for _ in range(8):
        opt.zero_grad()
        poses, scores = m(x)
        loss, _ = metrics.losses.training_loss(poses, scores, y)
        loss.backward()
        assert all(torch.isfinite(p.grad).all() for p in m.parameters())
        opt.step()
        losses.append(float(loss))
    print("losses", [round(v, 4) for v in losses])
    assert losses[-1] < losses[0]


def test_captured_train_step_matches_eager_steps():
    """The CUDA-graph replay of the whole training step performs the same updates as eager steps (no stochastic depth: identical
    arithmetic), including the device-side Adam step counter."""
    import manipose_b200 as mb
    from manipose_b200 import metrics
    from manipose_b200.optim import FusedAdam, CapturedTrainStep
    gen = torch.Generator().manual_seed(3)
    xs = [(0.3 * torch.randn(4, 27, 17, 2, generator=gen)).cuda() for _ in range(3)]
    ys = [(0.3 * torch.randn(4, 27, 17, 3, generator=gen)).cuda() for _ in range(3)]
    loss_fn = lambda out, y: metrics.losses.training_loss(out[0], out[1], y)[0]

    def make():
        torch.manual_seed(0)
        m = mb.RMCLManifoldMixSTE(mb.h36m17_skeleton(), num_frame=27, n_hyp=5, drop_path_rate=0.0).cuda().train()
        return m, FusedAdam(m, lr=1e-4, weight_decay=1e-6)

    m1, o1 = make()
    sd0 = {k: v.clone() for k, v in m1.state_dict().items()}
    eager, eager_m1 = [], None
    for x, y in zip(xs, ys):
        o1.zero_grad()
        loss = loss_fn(m1(x), y)
        loss.backward()
        o1.step()
        eager.append(float(loss.detach()))
        if eager_m1 is None:
            eager_m1 = (o1.exp_avg.clone(), o1.exp_avg_sq.clone())       # moments after ONE step: same parameters, same arithmetic
    m2, o2 = make()
    step = CapturedTrainStep(m2, o2, loss_fn, xs[0], ys[0], warmup=2)
    # the warm-up / capture steps leave no trace: parameters, moments and the step counter are what they were before
    for k, v in m2.state_dict().items():
        assert torch.equal(v, sd0[k]), k
    assert int(o2.step_dev) == 0 and o2.step_count == 0
    assert float(o2.exp_avg.abs().max()) == 0.0 and float(o2.exp_avg_sq.abs().max()) == 0.0
    graphed, graph_m1 = [], None
    for x, y in zip(xs, ys):
        graphed.append(float(step(x, y)))
        if graph_m1 is None:
            graph_m1 = (o2.exp_avg.clone(), o2.exp_avg_sq.clone())
    torch.cuda.synchronize()
    assert int(o2.step_dev) == 3
    # after the first step the two runs differ only by the order of fp32 atomics: the moments agree tightly
    assert _rel(graph_m1[0], eager_m1[0]) <= 2e-3 and _rel(graph_m1[1], eager_m1[1]) <= 2e-3
    assert abs(eager[0] - graphed[0]) <= 1e-5 * abs(eager[0])     # same parameters, same arithmetic
    for a, b in zip(eager, graphed):                                 # later steps: noise-signed first Adam updates (see below)
        assert abs(a - b) <= 2e-3 * abs(a), (eager, graphed)
    # Adam's first steps move every element by ~lr * sign(g): elements with |g| ~ 0 take the sign of fp32 reduction noise, so the
    # moments (linear / quadratic in the gradients), not the parameters, are what must agree
    # (and after the first step those noise-signed updates perturb the next gradients at the 1e-3 level)
    r1, r2 = _rel(o2.exp_avg, o1.exp_avg), _rel(o2.exp_avg_sq, o1.exp_avg_sq)
    print(f"captured vs eager after 3 steps: exp_avg rel-L2 {r1:.2e}, exp_avg_sq rel-L2 {r2:.2e}")
    assert r1 <= 1.5e-1 and r2 <= 5e-2          # measured 2e-2 .. 4e-2 / 6e-3: winner flips after the noise-signed first updates (chaotic)
    moved = float((m2.rotations_module.STEblocks[0].attn.qkv.weight.detach() - sd0["rotations_module.STEblocks.0.attn.qkv.weight"].cuda()).abs().mean())
    assert 0.5e-4 <= moved <= 4e-4             # three steps of ~lr = 1e-4 each
    # the learning rate is read from the device on replay: a scheduler's change takes effect (lr = 0 freezes the parameters exactly)
    before = o2.flat.flat_param.clone()
    o2.param_groups[0]["lr"] = 0.0
    step(xs[0], ys[0])
    torch.cuda.synchronize()
    assert torch.equal(o2.flat.flat_param, before)
    assert int(o2.step_dev) == 4
    o2.param_groups[0]["lr"] = 1e-4
    step(xs[0], ys[0])
    torch.cuda.synchronize()
    assert not torch.equal(o2.flat.flat_param, before)


@pytest.mark.parametrize("dtype", ["bf16", "fp16"])
def test_heads_training_kernels_vs_torch(dtype):
    """FoldedHeadsFn (mp_heads_fold -> mp_heads_fwd16; mp_heads_bwd_pack -> wgrad / dgrad -> mp_heads_unfold) against the K separate
    LayerNorm + Linear + score heads in fp32 torch on the same normalised 16-bit input: outputs, input gradient, every parameter gradient."""
    import manipose_b200 as mb
    from manipose_b200 import train_ops as T
    td = DT[dtype]
    torch.manual_seed(3)
    model = mb.RMCLManifoldMixSTE(mb.h36m17_skeleton(), num_frame=9, n_hyp=5).cuda().set_compute_dtype(dtype)
    heads = list(model.rotations_module.head)
    gen = torch.Generator(device="cuda").manual_seed(5)
    with torch.no_grad():
        for h in heads:
            for p in h.parameters():
                p.add_(0.05 * torch.randn(p.shape, generator=gen, device="cuda"))
    b, l, j, c, d = 3, 9, 17, 512, model.rotations_module.out_dim
    yhat = torch.randn(b * l * j, c, generator=gen, device="cuda").to(td)
    g_rot = torch.randn(b, 5, l, j, d, generator=gen, device="cuda")
    g_log = torch.randn(b, 5, l, generator=gen, device="cuda")
    # fp32 reference: LN_k(y) = yhat * gamma_k + beta_k (yhat already normalised), prediction Linear, score Linear over the joints
    yref = yhat.float().requires_grad_()
    rots, logs = [], []
    for h in heads:
        z = yref * h.norm.weight + h.norm.bias
        out = (z @ h.prediction_head.weight.t() + h.prediction_head.bias).reshape(b, l, j, d + 1)
        rots.append(out[..., :d])
        logs.append((out[..., d] * h.score_head.weight.reshape(1, 1, j)).sum(-1) + h.score_head.bias)
    rot_ref, log_ref = torch.stack(rots, 1), torch.stack(logs, 1)
    params = [p for h in heads for p in h.parameters()]
    ref_grads = torch.autograd.grad([rot_ref, log_ref], [yref] + params, [g_rot, g_log])
    for p in params:
        p.grad = None
    yin = yhat.clone().requires_grad_()
    rot, logits = T.FoldedHeadsFn.apply(yin, heads, b, l, j, d, heads[0].norm.weight)
    torch.autograd.backward([rot, logits], [g_rot, g_log])
    tol = RTOL[dtype]
    torch.testing.assert_close(rot, rot_ref.detach(), rtol=4 * tol, atol=4 * tol)
    torch.testing.assert_close(logits, log_ref.detach(), rtol=4 * tol, atol=8 * tol)
    rel = lambda a, r: float((a.float() - r).norm() / r.norm().clamp_min(1e-20))
    assert rel(yin.grad, ref_grads[0]) <= 3 * tol
    for p, gr in zip(params, ref_grads[1:]):
        assert rel(p.grad, gr) <= 3 * tol, (tuple(p.shape), rel(p.grad, gr))


def test_backward_streams_do_not_change_the_gradients():
    """Weight gradients on the second stream / the bone-length backbone on its own stream: same kernels, same inputs, so the same
    gradients as the in-line order (up to the order of the fp32 atomics of the bias / LayerNorm parameter sums)."""
    import manipose_b200 as mb
    from manipose_b200 import metrics
    from manipose_b200.architectures import mix_ste
    torch.manual_seed(0)
    model = mb.RMCLManifoldMixSTE(mb.h36m17_skeleton(), num_frame=27, n_hyp=5, drop_path_rate=0.0).cuda().train().set_compute_dtype("bf16")
    gen = torch.Generator().manual_seed(1)
    x = (0.3 * torch.randn(6, 27, 17, 2, generator=gen)).cuda()
    y = (0.3 * torch.randn(6, 27, 17, 3, generator=gen)).cuda()

    def grads(overlap):
        mix_ste.MixSTE.overlap_wgrad = overlap
        type(model).overlap_branches = overlap
        for p in model.parameters():
            p.grad = None
        loss, _ = metrics.losses.training_loss(*model(x), y)
        loss.backward()
        torch.cuda.synchronize()
        return float(loss.detach()), {n: p.grad.clone() for n, p in model.named_parameters()}

    was = (mix_ste.MixSTE.overlap_wgrad, type(model).overlap_branches)
    try:
        l0, g0 = grads(False)
        l1, g1 = grads(True)
        l2, g2 = grads(True)
    finally:
        mix_ste.MixSTE.overlap_wgrad, type(model).overlap_branches = was
    assert l0 == l1 == l2
    for n in g0:
        for g in (g1, g2):
            err = float((g[n] - g0[n]).norm() / g0[n].norm().clamp_min(1e-20))
            assert err <= 1e-5, (n, err)
