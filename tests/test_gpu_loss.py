"""Parity of the loss / hypothesis-metric kernels (through the C ABI) against the CPU oracle and the reference fixtures.
Winner and arg-max indices must be bit-exact on identical fp32 inputs (BASELINE.json north_star)."""
import os

import pytest
import torch

from oracle import manipose_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
W = O.STANDARD_H36M_WEIGHTS


def _load(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def test_golden_wta_and_scoring():
    import manipose_b200 as mb
    M = mb.metrics
    g = _load("loss.pt")
    hyp, scores, y = g["hyp"].cuda(), g["scores"].cuda(), g["y"].cuda()
    for name, w, sq in (("w", W, False), ("u", None, False), ("wsq", W, True)):
        v, i = M.wta_l2_loss_and_activate_head(hyp, y, w, sq)
        assert torch.equal(i.cpu(), g[f"wta_idx_{name}"]), "winner indices must be bit-exact"
        assert i.dtype == torch.int64
        torch.testing.assert_close(v.cpu(), g[f"wta_val_{name}"], rtol=1e-6, atol=1e-9)
        tot, bce = M.wta_with_scoring_loss(hyp, scores, y, 0.1, w, sq)
        torch.testing.assert_close(tot.cpu(), g[f"score_total_{name}"], rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(bce.cpu(), g[f"score_bce_{name}"], rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(M.mean_velocity_error(hyp, y, axis=2).cpu(), g["vel"], rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(M.mean_velocity_error(hyp, y, axis=2, squared=True).cpu(), g["vel_sq"], rtol=1e-5, atol=1e-8)
    torch.testing.assert_close(M.smoothness_regularization(hyp, W, axis=2).cpu(), g["smooth_w"], rtol=1e-5, atol=1e-8)
    assert M.wta_with_scoring_loss(hyp, scores, y, 0.0, W).dim() == 0   # reference quirk: bare scalar when beta == 0


def test_golden_training_loss_and_gradients():
    """make_loss + compute_and_acc_loss with config.yaml defaults (hpe/main_h36m_lifting.py:101-209), fused and as the four
    separate closures the driver builds; gradients w.r.t. poses and score logits vs the reference's autograd."""
    import manipose_b200 as mb
    from manipose_b200 import ops
    M = mb.metrics
    g = _load("loss.pt")
    y = g["y"].cuda()
    for fused in (True, False):
        hyp = g["hyp"].cuda().requires_grad_()
        logits = g["logits"].cuda().requires_grad_()
        scores = ops.softmax_hyp(logits)
        if fused:
            total, terms = M.training_loss(hyp, scores, y)
            torch.testing.assert_close(terms[0].cpu(), g["train_wloss"], rtol=1e-5, atol=1e-7)
            torch.testing.assert_close(0.1 * terms[1].cpu(), g["train_score_reg"], rtol=1e-5, atol=1e-7)
            torch.testing.assert_close(2.0 * terms[2].cpu(), g["train_vloss"], rtol=1e-5, atol=1e-7)
            torch.testing.assert_close(0.5 * terms[3].cpu(), g["train_sreg"], rtol=1e-5, atol=1e-7)
        else:
            wl = M.wta_l2_loss_and_activate_head(hypothesis=hyp, y=y, weights=W, squared=False)[0].mean()
            sr = M.wta_with_scoring_loss(hypothesis=hyp, scores=scores, y=y, beta=0.1, weights=W, squared=False)[1]
            vl = 2.0 * M.mean_velocity_error(predicted=hyp, target=y, squared=False, axis=2)
            sg = 0.5 * M.smoothness_regularization(prediction=hyp, weights=W, axis=2)
            total = wl + sr + vl + sg
        torch.testing.assert_close(total.cpu().reshape(1), g["train_total"], rtol=1e-5, atol=1e-7)
        total.backward()
        torch.testing.assert_close(hyp.grad.cpu(), g["grad_hyp"], rtol=1e-4, atol=1e-8)
        torch.testing.assert_close(logits.grad.cpu(), g["grad_logits"], rtol=1e-4, atol=1e-8)


def test_golden_aggregate_and_mpjpe():
    import manipose_b200 as mb
    g = _load("loss.pt")
    m = mb.RMCLManifoldMixSTE(mb.h36m17_skeleton(), num_frame=9, n_hyp=5, depth_rot=1, depth_seg=1)
    hyp, scores, y = g["hyp"].cuda(), g["scores"].cuda(), g["y"].cuda()
    torch.testing.assert_close(m.aggregate(hyp, scores, "weighted_ave").cpu(), g["agg_weighted"], rtol=1e-6, atol=1e-8)
    assert torch.equal(m.aggregate(hyp, scores, "best_score").cpu(), g["agg_best"])
    val, pose = m.aggregate(hyp, mode="oracle", ground_truth=y)
    torch.testing.assert_close(val.cpu(), g["agg_oracle_val"], rtol=1e-6, atol=1e-9)
    assert torch.equal(pose.cpu(), g["agg_oracle_pose"])
    idx = torch.argmax(scores, dim=1)[..., 0]
    assert torch.equal(m.poses_from_hyp_idx(hyp, idx).cpu(), g["agg_best"])
    assert m.concat_hyp_and_scores(hyp, scores).shape == (3, 5, 11, 17, 4)
    torch.testing.assert_close(mb.metrics.mpjpe_error(g["agg_weighted"].cuda(), y, "sum").cpu(), g["mpjpe_sum"], rtol=1e-6, atol=0)
    torch.testing.assert_close(mb.metrics.mpjpe_error(g["agg_weighted"].cuda(), y, "average").cpu(), g["mpjpe_avg"], rtol=1e-6, atol=0)
    with pytest.raises(AssertionError, match="Scores required"):
        m.aggregate(hyp, None, "weighted_ave")
    with pytest.raises(ValueError, match="Only best_score and weighted_ave"):
        m.aggregate(hyp, scores, "median")


@pytest.mark.parametrize("b,k,t", [(1, 1, 1), (2, 5, 27), (3, 10, 81), (30, 5, 27), (3, 5, 243), (2, 2, 33)])
def test_random_vs_oracle(b, k, t):
    import manipose_b200 as mb
    from manipose_b200 import ops
    M = mb.metrics
    gen = torch.Generator().manual_seed(100 * b + t)
    y = 0.3 * torch.randn(b, t, 17, 3, generator=gen)
    y[:, :, 0] = 0
    hyp = y[:, None] + 0.1 * torch.randn(b, k, t, 17, 3, generator=gen)
    if k > 1:
        hyp[:, 1] = hyp[:, 0]                     # exact ties between hypotheses 0 and 1 -> lowest index must win
    logits = torch.randn(b, k, t, 1, generator=gen)
    scores = logits.softmax(dim=1)
    for w, sq in ((W, False), (None, False), (W, True)):
        v_ref, i_ref = O.wta_l2_loss_and_activate_head(hyp, y, w, sq)
        v, i = M.wta_l2_loss_and_activate_head(hyp.cuda(), y.cuda(), w, sq)
        assert torch.equal(i.cpu(), i_ref)
        torch.testing.assert_close(v.cpu(), v_ref, rtol=1e-6, atol=1e-8)
    if t > 1:
        h_ref = hyp.clone().requires_grad_()
        l_ref = logits.clone().requires_grad_()
        tot_ref, _ = O.training_loss(h_ref, l_ref.softmax(dim=1), y)
        tot_ref.backward()
        h, lg = hyp.cuda().requires_grad_(), logits.cuda().requires_grad_()
        tot, _ = M.training_loss(h, ops.softmax_hyp(lg), y.cuda())
        tot.backward()
        torch.testing.assert_close(tot.cpu(), tot_ref, rtol=1e-5, atol=1e-7)
        # ties make the reference's min-backward pick hypothesis 0 too (torch.min(dim) routes to the returned index)
        torch.testing.assert_close(h.grad.cpu(), h_ref.grad, rtol=1e-4, atol=1e-8)
        torch.testing.assert_close(lg.grad.cpu(), l_ref.grad, rtol=1e-4, atol=1e-8)
    agg = ops.aggregate(hyp.cuda(), scores.reshape(b, k, t).cuda(), None, 0)[0]
    torch.testing.assert_close(agg.cpu(), O.aggregate(hyp, scores, "weighted_ave"), rtol=1e-6, atol=1e-8)
    best = ops.aggregate(hyp.cuda(), scores.reshape(b, k, t).cuda(), None, 1)
    assert torch.equal(best[2].cpu(), scores[..., 0].argmax(1))
    assert torch.equal(best[0].cpu(), O.aggregate(hyp, scores, "best_score"))


def test_error_paths_match_the_reference():
    import manipose_b200 as mb
    M = mb.metrics
    hyp = torch.zeros(1, 2, 4, 17, 3, device="cuda")
    y = torch.zeros(1, 4, 17, 3, device="cuda")
    with pytest.raises(AssertionError):
        M.wta_l2_loss_and_activate_head(hyp, y, torch.ones(15))       # losses.py:23 assert
    with pytest.raises(ValueError):
        M.wta_l2_loss_and_activate_head(hyp, y, None, True)             # reference: torch.min(dim=1) on a 0-d tensor raises
    with pytest.raises(ValueError):
        M.mpjpe_error(y, y, "median")


def test_full_size_loss_properties():
    """B=1024, T=243, K=5: (a) the oracle hypothesis has zero WTA loss and is picked, (b) loss is invariant to permuting
    clips, (c) a slice equals the oracle."""
    import manipose_b200 as mb
    M = mb.metrics
    b, k, t = 1024, 5, 243
    gen = torch.Generator(device="cuda").manual_seed(3)
    y = 0.3 * torch.randn(b, t, 17, 3, generator=gen, device="cuda")
    hyp = y[:, None] + 0.1 * torch.randn(b, k, t, 17, 3, generator=gen, device="cuda")
    winner = torch.randint(0, k, (b, t), generator=gen, device="cuda")
    hyp.scatter_(1, winner[:, None, :, None, None].expand(b, 1, t, 17, 3), y[:, None])
    val, idx = M.wta_l2_loss_and_activate_head(hyp, y, W)
    assert torch.equal(idx, winner) and bool((val == 0).all())
    scores = torch.softmax(torch.randn(b, k, t, 1, generator=gen, device="cuda"), dim=1)
    tot, terms = M.training_loss(hyp, scores, y)
    perm = torch.randperm(b, device="cuda")
    tot_p, _ = M.training_loss(hyp[perm].contiguous(), scores[perm].contiguous(), y[perm].contiguous())
    torch.testing.assert_close(tot, tot_p, rtol=1e-6, atol=0)
    sl = slice(100, 104)
    ref_tot, _ = O.training_loss(hyp[sl].cpu(), scores[sl].cpu(), y[sl].cpu())
    got_tot, _ = M.training_loss(hyp[sl].contiguous(), scores[sl].contiguous(), y[sl].contiguous())
    torch.testing.assert_close(got_tot.cpu(), ref_tot, rtol=1e-5, atol=1e-7)


# ------------------------------------------------------------------------------------------------ joint-error analytics
@pytest.mark.parametrize("shape", [(5, 27, 17, 3), (1, 1, 17, 3), (64, 243, 17, 3)])
def test_joint_error_analytics_vs_oracle(shape):
    """mse_error / jointwise_error / jointwise_mse / coordwise_error / segments_len_err / mpjpe_error(no_agg) through mp_point_errors,
    in millimetres like the drivers (main_h36m_lifting.py:975-1057), inputs on the CPU (moved by the functions) or on the device."""
    import manipose_b200 as mb
    from manipose_b200 import metrics as M
    sk = mb.h36m17_skeleton()
    gen = torch.Generator().manual_seed(sum(shape))
    pred = 300.0 * torch.randn(*shape, generator=gen)
    gt = 300.0 * torch.randn(*shape, generator=gen)
    pc, gc = pred.cuda(), gt.cuda()
    for mode in ("average", "sum"):
        rt = dict(rtol=2e-6, atol=0.0)       # fp64 partial sums here, fp32 pairwise sums in torch: agreement to fp32 rounding of the result
        torch.testing.assert_close(M.mse_error(pc, gc, mode).cpu(), O.mse_error(pred, gt, mode), **rt)
        torch.testing.assert_close(M.jointwise_error(pc, gc, mode).cpu(), O.jointwise_error(pred, gt, mode), **rt)
        torch.testing.assert_close(M.jointwise_mse(pc, gc, mode).cpu(), O.jointwise_error(pred, gt, mode, squared=True), **rt)
        torch.testing.assert_close(M.coordwise_error(pc, gc, mode).cpu(), O.coordwise_error(pred, gt, mode), **rt)
        for signed in (True, False):
            got = M.segments_len_err(batch_imp=pc.permute(0, 3, 2, 1), batch_gt=gc.permute(0, 3, 2, 1), skeleton=sk, mode=mode, signed=signed)
            want = O.segments_len_err(pred.permute(0, 3, 2, 1), gt.permute(0, 3, 2, 1), mode, signed)
            torch.testing.assert_close(got.cpu(), want, rtol=2e-5, atol=1e-3 if signed else 0.0)   # the signed sum cancels: absolute
    # element-wise modes: the same fp32 operations per element -> equal up to sqrt rounding (1 ulp)
    torch.testing.assert_close(M.mpjpe_error(pc, gc, "no_agg").cpu(), O.mpjpe_error(pred, gt, "no_agg"), rtol=2e-7, atol=0.0)
    assert torch.equal(M.mse_error(pc, gc, "no_agg").cpu(), O.mse_error(pred, gt, "no_agg"))
    torch.testing.assert_close(M.jointwise_error(pc, gc, "no_agg").cpu(), O.jointwise_error(pred, gt, "no_agg"), rtol=2e-7, atol=0.0)
    assert torch.equal(M.coordwise_error(pc, gc, "no_agg").cpu(), O.coordwise_error(pred, gt, "no_agg"))
    got = M.segments_len_err(batch_imp=pc.permute(0, 3, 2, 1), batch_gt=gc.permute(0, 3, 2, 1), skeleton=sk, mode="no_agg")
    # two bone lengths of ~500 mm each within 1 ulp (6e-5) of the reference's: their difference within a few ulp
    torch.testing.assert_close(got.cpu(), O.segments_len_err(pred.permute(0, 3, 2, 1), gt.permute(0, 3, 2, 1), "no_agg"), rtol=0.0, atol=3e-4)
    # CPU tensors and numpy arrays are accepted (moved to the device), like the reference's callers pass them
    torch.testing.assert_close(M.jointwise_error(pred, gt.numpy(), "average").cpu(), O.jointwise_error(pred, gt, "average"), rtol=2e-6, atol=0.0)
    assert abs(M.keypoint_3d_pck(pred.numpy().reshape(-1, 17, 3), gt.numpy().reshape(-1, 17, 3)) - O.keypoint_3d_pck(pred.reshape(-1, 17, 3), gt.reshape(-1, 17, 3))) <= 1e-4
    with pytest.raises(ValueError):
        M.jointwise_error(pc, gc, "median")


def test_weighted_losses_without_weights_and_over_all_hypotheses():
    """Reference behaviours that used to raise here: weighted_mse_loss(weights=None) = F.mse_loss (losses.py:57-58) and
    weighted_mpjpe_loss over K > 1 hypotheses with gradients (losses.py:14-43, every hypothesis contributes)."""
    import torch.nn.functional as F
    from manipose_b200 import metrics as M
    gen = torch.Generator().manual_seed(3)
    hyp = (0.3 * torch.randn(3, 4, 9, 17, 3, generator=gen)).cuda()
    y = (0.3 * torch.randn(3, 9, 17, 3, generator=gen)).cuda()
    w = M.STANDARD_H36M_WEIGHTS
    one = hyp[:, 0].clone().requires_grad_()
    got = M.weighted_mse_loss(one, y)
    ref_in = hyp[:, 0].clone().requires_grad_()
    want = F.mse_loss(ref_in, y)
    torch.testing.assert_close(got, want, rtol=1e-5, atol=0.0)
    got.backward()
    want.backward()
    torch.testing.assert_close(one.grad, ref_in.grad, rtol=1e-4, atol=1e-9)
    h1 = hyp.clone().requires_grad_()
    got = M.weighted_mpjpe_loss(h1, y[:, None].expand_as(hyp), w)
    h2 = hyp.clone().requires_grad_()
    want = torch.mean(w.cuda()[None, None, None, :] * torch.norm(h2 - y[:, None], p=2, dim=-1))
    torch.testing.assert_close(got, want, rtol=1e-5, atol=0.0)
    got.backward()
    want.backward()
    torch.testing.assert_close(h1.grad, h2.grad, rtol=1e-4, atol=1e-9)


# ------------------------------------------------------------------------------------------------ pose consistency
def _consistency_checks(poses_cpu, want, rtol=2e-5):
    import manipose_b200 as mb
    from manipose_b200 import metrics as M
    sk = mb.h36m17_skeleton()
    jc = poses_cpu.cuda().permute(0, 3, 2, 1)
    # correctly rounded ops in the reference's order; torch's CPU sqrt is itself 1 ulp off on ~0.7 % of inputs
    torch.testing.assert_close(M.measure_bones_length(jc, sk.bones).cpu(), want["bone_len"], rtol=3e-7, atol=0)
    for mode in ("average", "sum", "std", "min", "max"):
        torch.testing.assert_close(M.segments_time_consistency(jc, sk, mode).cpu(), want[f"stc_{mode}"], rtol=rtol, atol=1e-9)
    for mode in ("average", "sum", "std"):
        torch.testing.assert_close(M.segments_time_consistency_per_bone(jc, sk, mode).cpu(), want[f"stc_pb_{mode}"], rtol=rtol, atol=1e-9)
    for mode in ("average", "sum"):
        for squared in (True, False):
            torch.testing.assert_close(M.sagittal_symmetry(jc, sk, mode, squared).cpu(), want[f"sym_{mode}_{int(squared)}"], rtol=rtol, atol=1e-9)
            torch.testing.assert_close(M.sagittal_symmetry_per_bone(jc, sk, mode, squared).cpu(), want[f"sym_pb_{mode}_{int(squared)}"],
                                       rtol=rtol, atol=1e-9)


def test_pose_consistency_vs_reference_golden():
    """Bone lengths to 1 ulp, MPSCE / MPSSE statistics within fp32 reduction-order tolerance of the frozen reference outputs."""
    g = torch.load(os.path.join(GOLD, "consistency.pt"), weights_only=False)
    for tag, e in g.items():
        _consistency_checks(e["poses"], e)


@pytest.mark.parametrize("b,l", [(1024, 243), (1, 248832), (5, 1)])
def test_pose_consistency_vs_oracle_at_scale(b, l):
    """BASELINE config 3 size (1024 clips x 243 frames), the drivers' "all frames as one sequence" MPSCE (one clip of 248,832 frames,
    split over the SMs), and the single-frame edge (unbiased variance of one sample is NaN, like torch.var)."""
    from manipose_b200 import ops
    poses = 0.3 * torch.randn(b, l, 17, 3, generator=torch.Generator().manual_seed(b + l))
    jc = poses.permute(0, 3, 2, 1)
    lengths = O.measure_bones_length(jc)
    seg_mean, seg_var, sym_abs, sym_sq, bone_len = ops.pose_consistency(poses.cuda(), with_bone_lengths=True)
    torch.testing.assert_close(bone_len.cpu(), lengths, rtol=3e-7, atol=0)
    torch.testing.assert_close(seg_mean.cpu(), lengths.double().mean(2).float(), rtol=1e-6, atol=1e-7)
    if l == 1:
        assert torch.isnan(seg_var).all()
    else:
        torch.testing.assert_close(seg_var.cpu(), lengths.double().var(2).float(), rtol=1e-4, atol=1e-9)
    d = (lengths[:, list(O.H36M17_BONES_LEFT)] - lengths[:, list(O.H36M17_BONES_RIGHT)]).abs().double()
    torch.testing.assert_close(sym_abs.cpu(), d.mean(2).float(), rtol=1e-5, atol=1e-8)
    torch.testing.assert_close(sym_sq.cpu(), (d ** 2).mean(2).float(), rtol=1e-5, atol=1e-8)


# ------------------------------------------------------------------------------------------------ P-MPJPE
def test_p_mpjpe_vs_reference_golden():
    """MPJPE after per-frame Procrustes alignment vs the reference's numpy SVD implementation (noisy, similarity-transformed and
    mirrored predictions: the last one exercises the det(R) = -1 branch)."""
    from manipose_b200 import metrics as M
    g = torch.load(os.path.join(GOLD, "procrustes.pt"), weights_only=False)
    for name, e in g.items():
        if name == "pck":
            continue
        got = M.p_mpjpe(e["pred"].cuda(), e["target"].cuda())
        assert abs(got - e["p_mpjpe"]) <= 2e-5 * e["p_mpjpe"], (name, got, e["p_mpjpe"])


def test_pck_auc_vs_reference_golden_and_oracle():
    """3DPCK / AUC are counts: equal to the reference fixture (incl. errors exactly on thresholds) and to the numpy oracle at
    4.2 M points, up to the float32 rounding of count / N."""
    from manipose_b200 import metrics as M
    e = torch.load(os.path.join(GOLD, "procrustes.pt"), weights_only=False)["pck"]
    p, g = e["pred"].cuda(), e["gt"].cuda()
    assert abs(M.keypoint_3d_pck(p, g, threshold=150) - e["pck150"]) <= 1e-5
    assert abs(M.keypoint_3d_pck(p, g, threshold=50) - e["pck50"]) <= 1e-5
    assert abs(M.keypoint_3d_auc(p, g) - e["auc"]) <= 1e-5
    gen = torch.Generator().manual_seed(4)
    gt = 300.0 * torch.randn(248832, 17, 3, generator=gen)
    pred = gt + 70.0 * torch.randn(248832, 17, 3, generator=gen)
    assert abs(M.keypoint_3d_pck(pred.cuda(), gt.cuda()) - O.keypoint_3d_pck(pred, gt)) <= 1e-4
    assert abs(M.keypoint_3d_auc(pred.cuda(), gt.cuda()) - O.keypoint_3d_auc(pred, gt)) <= 1e-4
    with pytest.raises(NotImplementedError):
        M.keypoint_3d_pck(p, g, alignment="procrustes")
    with pytest.raises(ValueError):
        M.keypoint_3d_auc(p, g, alignment="affine")


def test_p_mpjpe_vs_oracle_at_scale_and_invariance():
    """62,208 frames (256 clips x 243) vs the numpy oracle, and the defining property: a similarity transform of the prediction
    leaves the aligned error unchanged (and a perfect prediction up to similarity scores ~0)."""
    from manipose_b200 import metrics as M
    gen = torch.Generator().manual_seed(9)
    y = 0.3 * torch.randn(256, 243, 17, 3, generator=gen)
    pred = y + 0.03 * torch.randn(256, 243, 17, 3, generator=gen)
    want = O.p_mpjpe(pred, y)
    got = M.p_mpjpe(pred.cuda(), y.cuda())
    assert abs(got - want) <= 2e-5 * want, (got, want)
    q, _ = torch.linalg.qr(torch.randn(3, 3, generator=gen))
    if torch.det(q) < 0:
        q[:, 0] *= -1
    moved = (0.6 * pred @ q + torch.tensor([1.0, 2.0, -0.5])).cuda()
    assert abs(M.p_mpjpe(moved, y.cuda()) - got) <= 1e-4 * got
    assert M.p_mpjpe((2.5 * y @ q + 0.7).cuda(), y.cuda()) <= 1e-5


@pytest.mark.parametrize("drop_last", [True, False])
def test_device_sequence_windows_vs_oracle(drop_last):
    """the device-side window gather returns exactly the items of the reference generator (restated in the oracle and
    pinned to it), including replicate-padded tails, ragged sequences shorter than one window, and arbitrary index order."""
    import numpy as np
    from manipose_b200.data import DeviceSequenceWindows
    rng = np.random.default_rng(1)
    lens = [243, 300, 27, 1000, 5, 486]
    p3 = [rng.standard_normal((n, 17, 3)).astype(np.float32) for n in lens]
    p2 = [rng.standard_normal((n, 17, 2)).astype(np.float32) for n in lens]
    for seq_len in (243, 27):
        items = O.sequence_windows(p3, p2, seq_len, drop_last)
        w = DeviceSequenceWindows(p3, p2, seq_len=seq_len, drop_last=drop_last)
        assert len(w) == len(items)
        order = list(reversed(range(len(items))))
        b2, b3 = w.batch(order)
        assert torch.equal(b2.cpu(), torch.stack([items[i][0] for i in order])) and torch.equal(b3.cpu(), torch.stack([items[i][1] for i in order]))
        got = [x for x in w.batches(4)]
        assert sum(x[0].shape[0] for x in got) == len(items)
        assert torch.equal(torch.cat([x[1] for x in got]).cpu(), torch.stack([it[1] for it in items]))


@pytest.mark.parametrize("miss_type", ["random", "random_left_arm_right_leg", "structured_joint", "structured_frame", "noisy", "all"])
@pytest.mark.parametrize("random_start", [False, True])
def test_device_sequence_windows_randomised_vs_oracle(miss_type, random_start):
    """SURVEY.md §8f-4, training-time side: random start frames, every occlusion pattern and the noisy input, sampled on the host with the
    reference's RNG calls and applied by the gather kernel, equal the oracle's items (pinned to the reference generator) under the same seeds."""
    import numpy as np
    from manipose_b200.data import DeviceSequenceWindows
    rng = np.random.default_rng(2)
    lens = [60, 300, 45, 100]
    p3 = [rng.standard_normal((n, 17, 3)).astype(np.float32) for n in lens]
    p2 = [rng.standard_normal((n, 17, 2)).astype(np.float32) for n in lens]
    order = [3, 0, 9, 1, 5, 8, 2]
    w = DeviceSequenceWindows(p3, p2, seq_len=27, drop_last=True, random_start=random_start, miss_type=miss_type, miss_rate=0.3, noise_sigma=0.05)
    torch.manual_seed(9)
    np.random.seed(9)
    items = O.sequence_windows(p3, p2, 27, True, random_start, miss_type, 0.3, 0.05, indices=order)
    torch.manual_seed(9)
    np.random.seed(9)
    b2, b3 = w.batch(order)
    assert b2.dtype == torch.float32 and b2.shape == (len(order), 27, 17, 2)
    assert torch.equal(b2.cpu(), torch.stack([it[0].float() for it in items]))      # "noisy" items are float64 in the reference; callers .float() them
    assert torch.equal(b3.cpu(), torch.stack([it[1] for it in items]))
    if miss_type not in ("noisy",):
        assert miss_type == "all" or float((b2 == 0).float().mean()) > 0.01           # something was actually occluded


@pytest.mark.parametrize("miss_type", ["no_miss", "random", "noisy"])
def test_device_sequence_windows_pose_flip_vs_oracle(miss_type):
    """The drivers' training loader (random starts + PoseFlip(0.5) + occlusion) through the device gather, same seeds as the oracle."""
    import numpy as np
    from manipose_b200.data import DeviceSequenceWindows
    rng = np.random.default_rng(5)
    lens = [60, 300, 45, 100]
    p3 = [rng.standard_normal((n, 17, 3)).astype(np.float32) for n in lens]
    p2 = [rng.standard_normal((n, 17, 2)).astype(np.float32) for n in lens]
    order = [3, 0, 9, 1, 5, 8, 2, 4, 6, 7]
    w = DeviceSequenceWindows(p3, p2, seq_len=27, drop_last=True, random_start=True, miss_type=miss_type, miss_rate=0.3, noise_sigma=0.05,
                              flip_probability=0.5)
    torch.manual_seed(12)
    np.random.seed(12)
    items = O.sequence_windows(p3, p2, 27, True, True, miss_type, 0.3, 0.05, indices=order, flip_probability=0.5)
    torch.manual_seed(12)
    np.random.seed(12)
    b2, b3 = w.batch(order)
    assert torch.equal(b2.cpu(), torch.stack([it[0].float() for it in items]))
    assert torch.equal(b3.cpu(), torch.stack([it[1] for it in items]))
    torch.manual_seed(12)
    np.random.seed(12)
    plain = O.sequence_windows(p3, p2, 27, True, True, "no_miss", 0.3, 0.05, indices=order, flip_probability=-1.0)   # same draws, never flips
    n_flipped = sum(int(not torch.equal(a[1], b[1])) for a, b in zip(items, plain))
    assert 0 < n_flipped < len(order)


def test_device_sequence_windows_vs_frozen_reference_items():
    """The device feed against items frozen from the UNMODIFIED reference generator (tests/golden/windows.pt): same seeds, same items
    (float32 view of them: the reference hands float64 sequences / noisy items to a caller that calls .float())."""
    import numpy as np
    from manipose_b200.data import DeviceSequenceWindows
    g = torch.load(os.path.join(GOLD, "windows.pt"), weights_only=False)
    for c in g["cases"]:
        p3, p2 = g["p3"][:c["n_seqs"]], g["p2"][:c["n_seqs"]]
        w = DeviceSequenceWindows(p3, p2, seq_len=g["seq_len"], drop_last=c["drop_last"], random_start=c["random_start"],
                                  miss_type=c["miss_type"], miss_rate=0.3, noise_sigma=0.05, flip_probability=c["flip"])
        assert len(w) == c["length"]
        torch.manual_seed(c["seed"])
        np.random.seed(c["seed"])
        b2, b3 = w.batch(c["order"])
        want2 = torch.stack([it[0].double() for it in c["items"]])
        want3 = torch.stack([it[1].float() for it in c["items"]])
        assert torch.equal(b3.cpu(), want3), c["miss_type"]
        # "noisy": the reference adds float64 noise to the float32 sequence and returns float64; the device rounds that sum once
        assert torch.equal(b2.cpu(), want2.float()), c["miss_type"]
