"""Checks oracle/manipose_oracle.py against fixtures frozen from the unmodified reference
(scripts/make_goldens.py).  Runs everywhere, including the GPU box where /root/reference is absent."""
import os

import pytest
import torch

from oracle import manipose_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def test_decoder_golden():
    g = _load("decoder.pt")
    n = g["rot6d"].shape[0]
    assert torch.equal(O.pose_decoder(g["rot6d"], g["bones"], torch.zeros(n, 3)), g["poses_zero_root"])
    assert torch.equal(O.pose_decoder(g["rot6d"], g["bones_signed"], g["root"]), g["poses_signed_root"])
    assert torch.equal(O.rotation_matrix_from_ortho6d(g["rot6d"].reshape(-1, 6)).reshape(n, 17, 3, 3), g["rotmats"])
    assert torch.equal(O.pose_decoder(g["rot4d"], g["bones"], torch.zeros(20, 3), rot_rep_dim=4), g["poses_4d"])


def test_decoder_ieee_restatement_is_pinned_to_the_reference():
    """pose_decoder_ieee (numpy, every op correctly rounded) vs the fixtures frozen from the reference: <= 1e-6 relative
    on well-conditioned rows (the residual is torch-CPU's 1-ulp sqrt), identical on the identity / T-pose known answer."""
    g = _load("decoder.pt")
    n = g["rot6d"].shape[0]
    ok = ~g["stress_rows"]
    for bones, root, key in ((g["bones"], torch.zeros(n, 3), "poses_zero_root"), (g["bones_signed"], g["root"], "poses_signed_root")):
        got, want = O.pose_decoder_ieee(g["rot6d"], bones, root), g[key]
        assert float((got[ok] - want[ok]).abs().max() / want[ok].abs().max()) <= 1e-6
        assert float((got[~ok] - want[~ok]).abs().max()) <= 1e-5
    ident = torch.tensor([1.0, 0, 0, 0, 1.0, 0]).expand(1, 17, 6).contiguous()
    assert torch.equal(O.pose_decoder_ieee(ident, g["kat_bones"], torch.zeros(1, 3)), g["kat_pose_identity"])
    got4 = O.pose_decoder_ieee(g["rot4d"], g["bones"], torch.zeros(20, 3))        # 4-D representation (rotation_tools.py:60-116)
    assert float((got4 - g["poses_4d"]).abs().max() / g["poses_4d"].abs().max()) <= 1e-6


def test_decoder_known_answers():
    """Identity 6-D reproduces the T-pose bit-exactly; joint values from SURVEY.md §8c (KAT with the
    bone lengths of hpe/useful_aux_scripts/test_forward_kinematics.py:104-106)."""
    g = _load("decoder.pt")
    ident = torch.tensor([1.0, 0, 0, 0, 1.0, 0]).expand(1, 17, 6).contiguous()
    pose = O.pose_decoder(ident, g["kat_bones"], torch.zeros(1, 3))
    assert torch.equal(pose, g["kat_pose_identity"])
    assert torch.equal(pose, g["kat_t_pose"])
    assert torch.equal(O.build_t_pose(g["kat_bones"]), g["kat_t_pose"])
    torch.testing.assert_close(pose[0, 3], torch.tensor([0.2, -1.0, 0.0]))
    torch.testing.assert_close(pose[0, 10], torch.tensor([0.0, 0.8, 0.0]))
    torch.testing.assert_close(pose[0, 13], torch.tensor([-1.0, 0.4, 0.0]))
    torch.testing.assert_close(pose[0, 16], torch.tensor([1.0, 0.4, 0.0]))


def test_decoder_properties():
    g = _load("decoder.pt")
    n = g["rot6d"].shape[0]
    poses = O.pose_decoder(g["rot6d"], g["bones"], torch.zeros(n, 3))
    assert torch.all(poses[:, 0] == 0)
    par = torch.tensor(O.H36M17_PARENTS[1:])
    lens = (poses[:, 1:] - poses[:, par]).norm(dim=-1)            # [N,16]
    ok = ~g["stress_rows"]
    per_clip = lens[ok].reshape(-1)  # bone lengths equal |bones| for well-conditioned rows
    want = g["bones"].abs().reshape(4, 1, 16).expand(4, n // 4, 16).reshape(n, 16)[ok].reshape(-1)
    torch.testing.assert_close(per_clip, want, rtol=1e-5, atol=1e-6)


def test_loss_golden():
    g = _load("loss.pt")
    hyp, scores, y = g["hyp"], g["scores"], g["y"]
    for name, w, sq in (("w", O.STANDARD_H36M_WEIGHTS, False), ("u", None, False), ("wsq", O.STANDARD_H36M_WEIGHTS, True)):
        v, i = O.wta_l2_loss_and_activate_head(hyp, y, w, sq)
        assert torch.equal(v, g[f"wta_val_{name}"]) and torch.equal(i, g[f"wta_idx_{name}"])
        tot, bce = O.wta_with_scoring_loss(hyp, scores, y, 0.1, w, sq)
        assert torch.equal(tot, g[f"score_total_{name}"]) and torch.equal(bce, g[f"score_bce_{name}"])
    assert torch.equal(O.mean_velocity_error(hyp, y, axis=2), g["vel"])
    assert torch.equal(O.mean_velocity_error(hyp, y, axis=2, squared=True), g["vel_sq"])
    assert torch.equal(O.smoothness_regularization(hyp, O.STANDARD_H36M_WEIGHTS, axis=2), g["smooth_w"])


def test_training_loss_and_gradients_golden():
    g = _load("loss.pt")
    hyp = g["hyp"].clone().requires_grad_(True)
    logits = g["logits"].clone().requires_grad_(True)
    tot, terms = O.training_loss(hyp, logits.softmax(dim=1), g["y"])
    torch.testing.assert_close(tot.reshape(1), g["train_total"], rtol=0, atol=1e-7)
    assert torch.equal(terms["wloss"], g["train_wloss"]) and torch.equal(terms["score_reg"], g["train_score_reg"])
    assert torch.equal(terms["vloss"], g["train_vloss"]) and torch.equal(terms["sreg"], g["train_sreg"])
    tot.backward()
    torch.testing.assert_close(hyp.grad, g["grad_hyp"], rtol=1e-6, atol=1e-9)
    torch.testing.assert_close(logits.grad, g["grad_logits"], rtol=1e-6, atol=1e-9)


def test_aggregate_golden():
    g = _load("loss.pt")
    hyp, scores, y = g["hyp"], g["scores"], g["y"]
    assert torch.equal(O.aggregate(hyp, scores, "weighted_ave"), g["agg_weighted"])
    assert torch.equal(O.aggregate(hyp, scores, "best_score"), g["agg_best"])
    v, p = O.aggregate(hyp, mode="oracle", ground_truth=y)
    assert torch.equal(v, g["agg_oracle_val"]) and torch.equal(p, g["agg_oracle_pose"])
    assert torch.equal(O.mpjpe_error(g["agg_weighted"], y, "sum"), g["mpjpe_sum"])
    assert torch.equal(O.mpjpe_error(g["agg_weighted"], y, "average"), g["mpjpe_avg"])


def test_forward_golden_synthetic_weights():
    """Seeded synthetic weights (regenerated here) pushed through the oracle must reproduce what the
    REFERENCE model produced with the same state_dict when the fixture was made."""
    g = _load("forward.pt")["t27k5_synth"]
    sd = O.make_state_dict(num_frame=g["T"], n_hyp=g["K"], seed=g["seed"])
    cs = (float(sum(t.double().sum() for t in sd.values())), float(sum(t.double().abs().sum() for t in sd.values())))
    if cs != tuple(g["checksum"]):
        pytest.skip("torch RNG stream differs from the build container: seeded weights are not the fixture's")
    with torch.no_grad():
        p, s = O.rmcl_forward(g["x"], sd)
    torch.testing.assert_close(s, g["scores"], rtol=0, atol=1e-6)
    torch.testing.assert_close(p, g["poses"], rtol=0, atol=2e-6)


def test_forward_golden_key_layout():
    g = _load("forward.pt")["t27k5_init"]
    want = dict(g["keys"])
    got = {n: tuple(t.shape) for n, t in O.make_state_dict(num_frame=27, n_hyp=5).items()}
    assert got == want and len(want) == 290


def test_pose_consistency_oracle_matches_frozen_reference_outputs():
    """bone lengths, MPSCE (segments_time_consistency) and MPSSE (sagittal_symmetry) frozen from the reference."""
    g = torch.load(os.path.join(GOLD, "consistency.pt"), weights_only=False)
    for tag, e in g.items():
        jc = e["poses"].permute(0, 3, 2, 1)
        assert torch.equal(O.measure_bones_length(jc), e["bone_len"]), tag
        for mode in ("average", "sum", "std", "min", "max"):
            assert torch.equal(O.segments_time_consistency(jc, mode), e[f"stc_{mode}"]), (tag, mode)
        for mode in ("average", "sum", "std"):
            assert torch.equal(O.segments_time_consistency(jc, mode, per_bone=True), e[f"stc_pb_{mode}"]), (tag, mode)
        for mode in ("average", "sum"):
            for squared in (True, False):
                assert torch.equal(O.sagittal_symmetry(jc, mode, squared), e[f"sym_{mode}_{int(squared)}"])
                assert torch.equal(O.sagittal_symmetry(jc, mode, squared, per_bone=True), e[f"sym_pb_{mode}_{int(squared)}"])


def test_tta_prediction_matches_frozen_reference_composition():
    """flip test-time augmentation (eval_utils.py:51-142) frozen from the reference's own model / pose_flip / aggregate."""
    g = torch.load(os.path.join(GOLD, "tta.pt"), weights_only=False)
    sd = O.make_state_dict(num_frame=g["T"], n_hyp=g["K"], seed=g["seed"])
    with torch.no_grad():
        for mode in ("weighted_ave", "best_score"):
            torch.testing.assert_close(O.tta_prediction(g["x"], sd, mode), g[mode], rtol=0, atol=1e-6)


def test_p_mpjpe_oracle_matches_frozen_reference_outputs():
    """Protocol #2 (MPJPE after Procrustes alignment) frozen from the reference's numpy implementation."""
    g = torch.load(os.path.join(GOLD, "procrustes.pt"), weights_only=False)
    for name, e in g.items():
        if name == "pck":
            assert O.keypoint_3d_pck(e["pred"], e["gt"], 150.0) == e["pck150"] and O.keypoint_3d_pck(e["pred"], e["gt"], 50.0) == e["pck50"]
            assert O.keypoint_3d_auc(e["pred"], e["gt"]) == e["auc"]
            continue
        assert abs(O.p_mpjpe(e["pred"], e["target"]) - e["p_mpjpe"]) <= 1e-6 * e["p_mpjpe"], name


def test_evaluate_oracle_matches_frozen_reference_outputs():
    """the whole of ``evaluate`` (hpe/eval_utils.py:16-203) frozen from the unmodified reference — predictions, MPJPE, the
    oracle and per-sample-oracle figures (with the reference's normalisation quirk), with and without flip TTA."""
    g = torch.load(os.path.join(GOLD, "evaluate.pt"), weights_only=False)
    sd = O.make_state_dict(num_frame=g["T"], n_hyp=g["K"], seed=g["seed"])
    for tta in (False, True):
        e = g[f"tta{int(tta)}"]
        with torch.no_grad():
            got = O.evaluate(g["batches"], sd, tta, return_hyps=False, compute_oracle=True)
        for a, b in zip(got[0] + got[5], e["predictions"] + e["oracle_preds"]):
            torch.testing.assert_close(a, b, rtol=0, atol=2e-3)          # mm
        assert abs(float(got[2]) - e["performance"]) <= 1e-3
        assert abs(float(got[3]) - e["oracle_mpjpe"]) <= 1e-3 and abs(float(got[4]) - e["psoracle_mpjpe"]) <= 1e-3


def test_sequence_windows_oracle_matches_frozen_reference_items():
    """items of the reference PoseSequenceGenerator under fixed torch / numpy seeds — replicate-padded tail, random starts,
    every occlusion pattern, the noisy input (float64 like the reference returns it) and the PoseFlip transform — bit for bit."""
    import numpy as np
    g = torch.load(os.path.join(GOLD, "windows.pt"), weights_only=False)
    assert len(g["cases"]) == 8
    for c in g["cases"]:
        p3, p2 = g["p3"][:c["n_seqs"]], g["p2"][:c["n_seqs"]]
        torch.manual_seed(c["seed"])
        np.random.seed(c["seed"])
        items = O.sequence_windows(p3, p2, g["seq_len"], c["drop_last"], c["random_start"], c["miss_type"], 0.3, 0.05, indices=c["order"],
                                   flip_probability=c["flip"])
        assert len(O.sequence_windows(p3, p2, g["seq_len"], c["drop_last"])) == c["length"]
        for (a2, a3), (r2, r3) in zip(items, c["items"]):
            assert a2.dtype == r2.dtype and torch.equal(a2, r2) and torch.equal(a3, r3), c["miss_type"]
