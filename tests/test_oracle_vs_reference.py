"""Pins oracle/manipose_oracle.py against the UNMODIFIED reference imported from /root/reference.

Skipped where the reference is absent (the GPU box); tests/test_oracle_golden.py covers that case
with frozen reference outputs."""
import pytest
import torch

from oracle import manipose_oracle as O
from oracle.ref_loader import load_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def ref():
    return load_reference()


def _perturb(model, seed=7, std=0.02):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in model.parameters():
            p.add_(torch.randn(p.shape, generator=g) * std)


@pytest.mark.parametrize("n_hyp,T,perturb", [(5, 27, False), (5, 27, True), (1, 9, True), (10, 9, True)])
def test_full_forward_matches_reference(ref, n_hyp, T, perturb):
    torch.manual_seed(42)
    m = ref.architectures.RMCLManifoldMixSTE(ref.make_skeleton(), num_frame=T, n_hyp=n_hyp,
                                             drop_path_rate=0.1).eval()
    if perturb:
        _perturb(m)
    x = 0.3 * torch.randn(2, T, 17, 2, generator=torch.Generator().manual_seed(1234))
    with torch.no_grad():
        p_ref, s_ref = m(x)
        p, s = O.rmcl_forward(x, m.state_dict())
    assert torch.equal(s, s_ref)
    torch.testing.assert_close(p, p_ref, rtol=0, atol=1e-6)


def test_single_hypothesis_model_matches_reference(ref):
    torch.manual_seed(3)
    m = ref.architectures.ManifoldMixSTE(ref.make_skeleton(), num_frame=9, drop_path_rate=0.1).eval()
    _perturb(m)
    x = 0.3 * torch.randn(2, 9, 17, 2, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        torch.testing.assert_close(O.manifold_forward(x, m.state_dict()), m(x), rtol=0, atol=1e-6)


def test_synthetic_state_dict_has_reference_keys_and_shapes(ref):
    for k in (1, 5):
        m = ref.architectures.RMCLManifoldMixSTE(ref.make_skeleton(), num_frame=27, n_hyp=k)
        want = {n: tuple(t.shape) for n, t in m.state_dict().items()}
        got = {n: tuple(t.shape) for n, t in O.make_state_dict(num_frame=27, n_hyp=k).items()}
        assert got == want
    m = ref.architectures.RMCLManifoldMixSTE(ref.make_skeleton(), num_frame=27, n_hyp=5)
    m.load_state_dict(O.make_state_dict(num_frame=27, n_hyp=5, seed=3))
    x = 0.3 * torch.randn(1, 27, 17, 2, generator=torch.Generator().manual_seed(9))
    with torch.no_grad():
        p_ref, s_ref = m.eval()(x)
    p, s = O.rmcl_forward(x, O.make_state_dict(num_frame=27, n_hyp=5, seed=3))
    assert torch.equal(s, s_ref)
    torch.testing.assert_close(p, p_ref, rtol=0, atol=1e-6)


def _decoder_inputs(n_clips=6, k=5, t=9, seed=0, stress=True, signed=False):
    g = torch.Generator().manual_seed(seed)
    n = n_clips * k * t
    r6 = torch.randn(n, 17, 6, generator=g)
    if stress:
        r6[::17, 3] *= 1e-9                       # tiny-norm rows -> the 1e-8 clamp
        r6[5::23, 7, 3:6] = 2.5 * r6[5::23, 7, 0:3]  # collinear a,b -> degenerate cross product
    bones = 0.1 + 0.4 * torch.rand(n_clips, 16, 1, generator=g)
    if signed:
        bones = bones * torch.where(torch.rand(n_clips, 16, 1, generator=g) < 0.3, -1.0, 1.0)
    return r6, bones


@pytest.mark.parametrize("signed", [False, True])
def test_decoder_bit_exact_vs_reference(ref, signed):
    r6, bones = _decoder_inputs(signed=signed)
    dec = ref.PoseDecoder(ref.make_skeleton(), rot_rep_dim=6)
    root = torch.randn(r6.shape[0], 3, generator=torch.Generator().manual_seed(1))
    assert torch.equal(O.pose_decoder(r6, bones, root), dec(r6, bones, root))
    assert torch.equal(O.rotation_matrix_from_ortho6d(r6.reshape(-1, 6)),
                       ref.rotation_tools.compute_rotation_matrix_from_ortho6d(r6.reshape(-1, 6)))


def test_decoder_4d_vs_reference(ref):
    g = torch.Generator().manual_seed(2)
    r4 = torch.randn(30, 17, 4, generator=g)
    bones = 0.1 + 0.4 * torch.rand(3, 16, 1, generator=g)
    dec = ref.PoseDecoder(ref.make_skeleton(), rot_rep_dim=4)
    root = torch.zeros(30, 3)
    assert torch.equal(O.pose_decoder(r4, bones, root, rot_rep_dim=4), dec(r4, bones, root))


def _loss_inputs(b=3, k=5, t=9, seed=0):
    g = torch.Generator().manual_seed(seed)
    y = 0.3 * torch.randn(b, t, 17, 3, generator=g)
    y[:, :, 0] = 0
    hyp = y[:, None] + 0.1 * torch.randn(b, k, t, 17, 3, generator=g)
    scores = torch.randn(b, k, t, 1, generator=g).softmax(dim=1)
    return hyp, scores, y


@pytest.mark.parametrize("weighted", [False, True])
@pytest.mark.parametrize("squared", [False, True])
def test_losses_bit_exact_vs_reference(ref, weighted, squared):
    M = ref.metrics
    hyp, scores, y = _loss_inputs()
    w_ref = M.STANDARD_H36M_WEIGHTS if weighted else None
    w = O.STANDARD_H36M_WEIGHTS if weighted else None
    if squared and not weighted:
        pytest.skip("reference returns a 0-d tensor and torch.min(dim=1) raises (losses.py:57-58)")
    v_ref, i_ref = M.wta_l2_loss_and_activate_head(hyp, y, w_ref, squared)
    v, i = O.wta_l2_loss_and_activate_head(hyp, y, w, squared)
    assert torch.equal(v, v_ref) and torch.equal(i, i_ref) and i.dtype == torch.int64
    tot_ref, bce_ref = M.wta_with_scoring_loss(hyp, scores, y, 0.1, w_ref, squared)
    tot, bce = O.wta_with_scoring_loss(hyp, scores, y, 0.1, w, squared)
    assert torch.equal(tot, tot_ref) and torch.equal(bce, bce_ref)
    assert torch.equal(O.wta_with_scoring_loss(hyp, scores, y, 0, w, squared),
                       M.wta_with_scoring_loss(hyp, scores, y, 0, w_ref, squared))
    assert torch.equal(O.mean_velocity_error(hyp, y, axis=2, squared=squared),
                       M.mean_velocity_error(hyp, y, axis=2, squared=squared))
    if weighted:
        assert torch.equal(O.smoothness_regularization(hyp, w, axis=2),
                           M.smoothness_regularization(hyp, w_ref, axis=2))
    else:  # reference quirk: weights=None only works on 4-D input (regularizations.py:165-170)
        assert torch.equal(O.smoothness_regularization(hyp[:, 0], None, axis=1),
                           M.smoothness_regularization(hyp[:, 0], None, axis=1))
        with pytest.raises(AssertionError):
            M.smoothness_regularization(hyp, None, axis=2)
        with pytest.raises(AssertionError):
            O.smoothness_regularization(hyp, None, axis=2)


def test_aggregate_and_metrics_vs_reference(ref):
    torch.manual_seed(0)
    m = ref.architectures.RMCLManifoldMixSTE(ref.make_skeleton(), num_frame=9, n_hyp=5)
    hyp, scores, y = _loss_inputs()
    assert torch.equal(O.aggregate(hyp, scores, "weighted_ave"), m.aggregate(hyp, scores, "weighted_ave"))
    assert torch.equal(O.aggregate(hyp, scores, "best_score"), m.aggregate(hyp, scores, "best_score"))
    v_ref, p_ref = m.aggregate(hyp, mode="oracle", ground_truth=y)
    v, p = O.aggregate(hyp, mode="oracle", ground_truth=y)
    assert torch.equal(v, v_ref) and torch.equal(p, p_ref)
    assert torch.equal(O.concat_hyp_and_scores(hyp, scores), m.concat_hyp_and_scores(hyp, scores))
    for mode in ("sum", "average"):
        assert torch.equal(O.mpjpe_error(hyp[:, 0], y, mode), ref.metrics.mpjpe_error(hyp[:, 0], y, mode))
    with pytest.raises(ValueError):
        O.aggregate(hyp, scores, "nope")


def test_training_loss_matches_reference_closures(ref):
    """make_loss/compute_and_acc_loss (hpe/main_h36m_lifting.py:101-209) restated with config.yaml defaults."""
    M = ref.metrics
    hyp, scores, y = _loss_inputs(seed=4)
    w = M.STANDARD_H36M_WEIGHTS
    wl = M.wta_l2_loss_and_activate_head(hypothesis=hyp, y=y, weights=w, squared=False)[0].mean()
    sr = M.wta_with_scoring_loss(hypothesis=hyp, scores=scores, y=y, beta=0.1, weights=w, squared=False)[1]
    vl = 2.0 * M.mean_velocity_error(predicted=hyp, target=y, squared=False, axis=2)
    sg = 0.5 * M.smoothness_regularization(prediction=hyp, weights=w, axis=2)
    loss = torch.zeros(1)
    for t in (wl, sr, vl, sg):
        loss += t
    tot, terms = O.training_loss(hyp, scores, y)
    torch.testing.assert_close(tot.reshape(1), loss, rtol=0, atol=1e-7)
    assert torch.equal(terms["wloss"], wl) and torch.equal(terms["score_reg"], sr)
    assert torch.equal(terms["vloss"], vl) and torch.equal(terms["sreg"], sg)


def test_pose_flip_vs_reference(ref):
    import importlib
    fn = importlib.import_module("mh_so3_hpe.augmentations.functional").pose_flip
    x = torch.randn(2, 5, 17, 3)
    sk = ref.make_skeleton()
    (want,) = fn((x.clone(),), sk)   # in-place on a tuple of tensors (functional.py:11-27)
    assert torch.equal(O.pose_flip(x), want)


def test_joint_error_analytics_match_reference(ref):
    """SURVEY.md §8f-3, the rest of mean_joint_errors.py:39-141: mse_error, jointwise_error, jointwise_mse, coordwise_error and
    segments_len_err in the layouts the drivers call them with (main_h36m_lifting.py:975-1057)."""
    sk = ref.make_skeleton()
    gen = torch.Generator().manual_seed(17)
    pred = 300.0 * torch.randn(5, 27, 17, 3, generator=gen)
    gt = 300.0 * torch.randn(5, 27, 17, 3, generator=gen)
    M = ref.metrics
    for mode in ("average", "sum", "no_agg"):
        assert torch.equal(O.mse_error(pred, gt, mode), M.mse_error(pred, gt, mode))
        assert torch.equal(O.jointwise_error(pred, gt, mode), M.jointwise_error(pred, gt, mode))
        assert torch.equal(O.jointwise_error(pred, gt, mode, squared=True), M.jointwise_mse(pred, gt, mode))
        assert torch.equal(O.coordwise_error(pred, gt, mode), M.coordwise_error(pred, gt, mode))
        for signed in (True, False):
            want = M.segments_len_err(batch_imp=pred.permute(0, 3, 2, 1), batch_gt=gt.permute(0, 3, 2, 1), skeleton=sk, mode=mode, signed=signed)
            assert torch.equal(O.segments_len_err(pred.permute(0, 3, 2, 1), gt.permute(0, 3, 2, 1), mode, signed), want)


@pytest.mark.parametrize("b,l", [(3, 27), (1, 243)])
def test_pose_consistency_metrics_match_reference(ref, b, l):
    """measure_bones_length / segments_time_consistency (MPSCE) / sagittal_symmetry (MPSSE) restatements."""
    sk = ref.make_skeleton()
    assert tuple((int(j), int(p)) for j, p in sk.bones) == O.H36M17_BONES
    assert tuple(sk.bones_left) == O.H36M17_BONES_LEFT and tuple(sk.bones_right) == O.H36M17_BONES_RIGHT
    poses = 0.3 * torch.randn(b, l, 17, 3, generator=torch.Generator().manual_seed(b + l))
    jc = poses.permute(0, 3, 2, 1)
    M = ref.metrics
    assert torch.equal(O.measure_bones_length(jc), M.measure_bones_length(jc, sk.bones))
    for mode in ("average", "sum", "std", "min", "max"):
        assert torch.equal(O.segments_time_consistency(jc, mode), M.segments_time_consistency(jc, sk, mode))
    for mode in ("average", "sum", "std"):
        assert torch.equal(O.segments_time_consistency(jc, mode, per_bone=True), M.segments_time_consistency_per_bone(jc, sk, mode))
    for mode in ("average", "sum"):
        for squared in (True, False):
            assert torch.equal(O.sagittal_symmetry(jc, mode, squared), M.sagittal_symmetry(jc, sk, mode, squared))
            assert torch.equal(O.sagittal_symmetry(jc, mode, squared, per_bone=True), M.sagittal_symmetry_per_bone(jc, sk, mode, squared))


@pytest.mark.parametrize("mode", ["weighted_ave", "best_score"])
def test_tta_prediction_matches_reference_composition(ref, mode):
    """the TTA branch of hpe/eval_utils.py:51-142 composed from the reference's model, pose_flip and aggregate."""
    from mh_so3_hpe.augmentations.functional import pose_flip
    sk = ref.make_skeleton()
    torch.manual_seed(11)
    m = ref.architectures.RMCLManifoldMixSTE(sk, num_frame=9, n_hyp=3, drop_path_rate=0.1).eval()
    _perturb(m)
    x = 0.3 * torch.randn(2, 9, 17, 2, generator=torch.Generator().manual_seed(8))
    with torch.no_grad():
        pred = m.aggregate(*m(x.clone()), mode=mode)
        pred_f = m.aggregate(*m(pose_flip(poses_tuple=(x.clone(),), skeleton=sk)[0]), mode=mode)
        want = (pred + pose_flip(poses_tuple=(pred_f,), skeleton=sk)[0]) / 2
        got = O.tta_prediction(x, m.state_dict(), mode)
    torch.testing.assert_close(got, want, rtol=0, atol=1e-6)


def _procrustes_cases():
    g = torch.Generator().manual_seed(17)
    y = 0.3 * torch.randn(3, 27, 17, 3, generator=g)
    y[:, :, 0] = 0
    cases = {"noisy": (y + 0.05 * torch.randn(3, 27, 17, 3, generator=g), y)}
    q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g))
    cases["similarity"] = (1.7 * y @ q + torch.tensor([0.3, -0.2, 0.9]) + 0.01 * torch.randn(3, 27, 17, 3, generator=g), y)
    mirrored = y.clone()
    mirrored[..., 0] *= -1                                  # a reflection: the det(R) = -1 branch
    cases["mirrored"] = (mirrored + 0.02 * torch.randn(3, 27, 17, 3, generator=g), y)
    return cases


def test_p_mpjpe_matches_reference(ref):
    """Protocol #2 restatement vs the reference's numpy implementation (same calls, so bit-equal)."""
    from mh_so3_hpe.metrics.mean_joint_errors import p_mpjpe
    for name, (pred, y) in _procrustes_cases().items():
        assert O.p_mpjpe(pred, y) == float(p_mpjpe(pred, y)), name


def test_pck_auc_match_reference(ref):
    """3DPCK / AUC restatements (alignment 'none', no mask) vs the reference's numpy functions."""
    from mh_so3_hpe.metrics.pck import keypoint_3d_pck, keypoint_3d_auc
    g = torch.Generator().manual_seed(23)
    gt = 300.0 * torch.randn(500, 17, 3, generator=g)                       # millimetres, like main_3dhp.py:880
    pred = gt + 60.0 * torch.randn(500, 17, 3, generator=g)
    for thr in (150.0, 50.0):
        assert O.keypoint_3d_pck(pred, gt, thr) == float(keypoint_3d_pck(pred, gt, threshold=thr))
    assert O.keypoint_3d_auc(pred, gt) == float(keypoint_3d_auc(pred, gt))


@pytest.mark.parametrize("drop_last", [True, False])
def test_sequence_windows_match_reference_generator(ref, drop_last):
    """clip windowing of PoseSequenceGenerator (fixed starts, no missing joints), incl. the replicate-padded tail."""
    import numpy as np
    from mh_so3_hpe.data.generators import PoseSequenceGenerator
    rng = np.random.default_rng(0)
    lens = [27, 40, 9, 81, 5]
    p3 = [rng.standard_normal((n, 17, 3)).astype(np.float32) for n in lens]
    p2 = [rng.standard_normal((n, 17, 2)).astype(np.float32) for n in lens]
    gen = PoseSequenceGenerator(p3, p2, None, seq_len=9, random_start=False, drop_last=drop_last, miss_type="no_miss")
    items = O.sequence_windows(p3, p2, 9, drop_last)
    assert len(items) == len(gen)
    for i, (a2, a3) in enumerate(items):
        r2, r3 = gen[i]
        assert torch.equal(a2, r2) and torch.equal(a3, r3), i


def _load_reference_evaluate():
    """hpe/eval_utils.py imports omegaconf only for a type hint (:8): stub it, then import the module from the reference tree."""
    import importlib.util
    import os
    import sys
    import types
    from oracle.ref_loader import REFERENCE_ROOT
    if "omegaconf" not in sys.modules:
        om = types.ModuleType("omegaconf")
        om.DictConfig = dict
        sys.modules["omegaconf"] = om
    spec = importlib.util.spec_from_file_location("_ref_eval_utils", os.path.join(REFERENCE_ROOT, "hpe", "eval_utils.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("tta", [False, True])
@pytest.mark.parametrize("return_hyps", [False, True])
def test_evaluate_matches_reference(ref, tta, return_hyps):
    """the whole of ``evaluate`` (hpe/eval_utils.py:16-203) — predictions, MPJPE, oracle and per-sample-oracle figures,
    with and without flip TTA — run unmodified on CPU against the oracle's restatement of its bookkeeping."""
    import types
    ev = _load_reference_evaluate()
    sk = ref.make_skeleton()
    torch.manual_seed(11)
    m = ref.architectures.RMCLManifoldMixSTE(sk, num_frame=9, n_hyp=3, drop_path_rate=0.1).eval()
    _perturb(m)
    g = torch.Generator().manual_seed(8)
    batches = [(0.3 * torch.randn(b, 9, 17, 2, generator=g), 0.3 * torch.randn(b, 9, 17, 3, generator=g)) for b in (2, 3)]
    cfg = types.SimpleNamespace(train=types.SimpleNamespace(tta=tta))
    want = ev.evaluate(m, [(x.clone(), y.clone()) for x, y in batches], "cpu", cfg, sk, return_hyps=return_hyps, compute_oracle=True)
    got = O.evaluate(batches, m.state_dict(), tta, return_hyps=return_hyps, compute_oracle=True)
    assert len(want) == len(got) == 6
    for a, b in zip(got[0] + got[1] + got[5], want[0] + want[1] + want[5]):
        torch.testing.assert_close(a, b, rtol=0, atol=2e-3)           # mm
    assert abs(float(got[2]) - float(want[2])) <= 1e-3
    assert abs(float(got[3]) - float(want[3])) <= 1e-3 and abs(float(got[4]) - float(want[4])) <= 1e-3
    want3 = ev.evaluate(m, [(x.clone(), y.clone()) for x, y in batches], "cpu", cfg, sk, return_hyps=return_hyps, compute_oracle=False)
    got3 = O.evaluate(batches, m.state_dict(), tta, return_hyps=return_hyps, compute_oracle=False)
    assert len(want3) == len(got3) == 3 and abs(float(got3[2]) - float(want3[2])) <= 1e-3


@pytest.mark.parametrize("miss_type", ["random", "random_left_arm_right_leg", "structured_joint", "structured_frame", "noisy", "all"])
def test_randomised_sequence_windows_match_reference_generator(ref, miss_type):
    """SURVEY.md §8f-4, training-time side of PoseSequenceGenerator: random start frames and every occlusion pattern, with torch's and
    numpy's global RNGs seeded identically, reproduce the reference's items bit for bit."""
    import numpy as np
    from mh_so3_hpe.data.generators import PoseSequenceGenerator
    rng = np.random.default_rng(3)
    lens = [60, 45, 100]
    p3 = [rng.standard_normal((n, 17, 3)) for n in lens]            # float64: the reference's in-place noise then cannot leak into the dataset
    p2 = [rng.standard_normal((n, 17, 2)) for n in lens]
    order = [4, 0, 7, 2, 5, 1]
    for random_start in (False, True):
        gen = PoseSequenceGenerator(p3, p2, None, seq_len=20, random_start=random_start, drop_last=True, miss_type=miss_type, miss_rate=0.3,
                                    noise_sigma=0.05)
        torch.manual_seed(5)
        np.random.seed(5)
        want = [gen[i] for i in order]
        torch.manual_seed(5)
        np.random.seed(5)
        got = O.sequence_windows(p3, p2, 20, True, random_start, miss_type, 0.3, 0.05, indices=order)
        for (a2, a3), (r2, r3) in zip(got, want):
            assert torch.equal(a2, r2) and torch.equal(a3, r3)


@pytest.mark.parametrize("miss_type", ["no_miss", "random", "all"])
def test_sequence_windows_with_pose_flip_transform_match_reference(ref, miss_type):
    """The training loader of the drivers (main_h36m_lifting.py:583-595): random starts + PoseFlip(probability 0.5) + occlusion pattern."""
    import numpy as np
    from mh_so3_hpe.data.generators import PoseSequenceGenerator
    from mh_so3_hpe.augmentations.transforms import PoseFlip
    rng = np.random.default_rng(4)
    lens = [60, 45, 100]
    p3 = [rng.standard_normal((n, 17, 3)) for n in lens]            # float64: the reference's in-place flip then cannot leak into the dataset
    p2 = [rng.standard_normal((n, 17, 2)) for n in lens]
    order = [4, 0, 7, 2, 5, 1, 3, 6]
    gen = PoseSequenceGenerator(p3, p2, None, seq_len=20, random_start=True, drop_last=True, miss_type=miss_type, miss_rate=0.3,
                                transform=PoseFlip(skeleton=ref.make_skeleton(), probability=0.5))
    torch.manual_seed(6)
    np.random.seed(6)
    want = [gen[i] for i in order]
    torch.manual_seed(6)
    np.random.seed(6)
    got = O.sequence_windows(p3, p2, 20, True, True, miss_type, 0.3, 5, indices=order, flip_probability=0.5)
    flipped = 0
    for (a2, a3), (r2, r3) in zip(got, want):
        assert torch.equal(a2, r2) and torch.equal(a3, r3)
    # some, not all, items were flipped (the coin is shared between the two poses of an item)
    torch.manual_seed(6)
    np.random.seed(6)
    plain = O.sequence_windows(p3, p2, 20, True, True, "no_miss", 0.3, 5, indices=order, flip_probability=-1.0)   # same draws, never flips
    flipped = sum(int(not torch.equal(a[1], b[1])) for a, b in zip(got, plain))
    assert 0 < flipped < len(order)


@pytest.mark.parametrize("return_hyps", [False, True])
def test_lift_action_postprocessing_matches_reference(ref, return_hyps):
    """``lift_action`` (hpe/eval_utils.py:226-251) = evaluate + a reshape to [N*L, ...] in metres: the product's ``stack_lifted`` applied to
    the oracle's evaluate output equals the unmodified reference's lift_action on the same model (CPU, no kernel involved)."""
    import types
    import numpy as np
    from manipose_b200.evaluation import stack_lifted
    ev = _load_reference_evaluate()
    sk = ref.make_skeleton()
    torch.manual_seed(11)
    m = ref.architectures.RMCLManifoldMixSTE(sk, num_frame=9, n_hyp=3, drop_path_rate=0.1).eval()
    _perturb(m)
    g = torch.Generator().manual_seed(8)
    batches = [(0.3 * torch.randn(b, 9, 17, 2, generator=g), 0.3 * torch.randn(b, 9, 17, 3, generator=g)) for b in (2, 3)]
    cfg = types.SimpleNamespace(train=types.SimpleNamespace(tta=False))
    want = ev.lift_action([(x.clone(), y.clone()) for x, y in batches], m, "cpu", cfg, sk, return_hyps)
    got = stack_lifted(O.evaluate(batches, m.state_dict(), False, return_hyps=return_hyps)[0])
    assert got.shape == want.shape == ((45, 3, 17, 4) if return_hyps else (45, 17, 3))
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-6)
