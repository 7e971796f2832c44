"""Freezes outputs of the UNMODIFIED reference (imported read-only from /root/reference through
oracle/ref_loader.py) into small fixtures under tests/golden/.  Run in the build container only:

    python scripts/make_goldens.py

The reference has no golden vectors of its own; these are the pins the GPU box uses,
because /root/reference does not exist there.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_loader import load_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def weights_checksum(sd):
    return float(sum(t.double().sum() for t in sd.values())), float(sum(t.double().abs().sum() for t in sd.values()))


def decoder_golden(ref):
    g = torch.Generator().manual_seed(1234)
    n_clips, k, t = 4, 5, 13
    n = n_clips * k * t
    r6 = torch.randn(n, 17, 6, generator=g)
    stress = torch.zeros(n, dtype=torch.bool)
    r6[::17, 3] *= 1e-9
    stress[::17] = True
    r6[5::23, 7, 3:6] = 2.5 * r6[5::23, 7, 0:3]
    stress[5::23] = True
    bones = 0.1 + 0.4 * torch.rand(n_clips, 16, 1, generator=g)
    bones_signed = bones * torch.where(torch.rand(n_clips, 16, 1, generator=g) < 0.3, -1.0, 1.0)
    root = torch.randn(n, 3, generator=g)
    dec = ref.PoseDecoder(ref.make_skeleton(), rot_rep_dim=6)
    out = {
        "rot6d": r6, "stress_rows": stress, "bones": bones, "bones_signed": bones_signed, "root": root,
        "poses_zero_root": dec(r6, bones, torch.zeros(n, 3)),
        "poses_signed_root": dec(r6, bones_signed, root),
        "rotmats": ref.rotation_tools.compute_rotation_matrix_from_ortho6d(r6.reshape(-1, 6)).reshape(n, 17, 3, 3),
        "K": k, "T": t,
    }
    # KAT from hpe/useful_aux_scripts/test_forward_kinematics.py:104-106 (bone lengths) with identity 6-D
    kat_len = torch.tensor([0.2, 0.5, 0.5, 0.2, 0.5, 0.5, 0.2, 0.2, 0.2, 0.2, 0.2, 0.4, 0.4, 0.2, 0.4, 0.4]).view(1, 16, 1)
    ident = torch.tensor([1.0, 0, 0, 0, 1.0, 0]).expand(1, 17, 6).contiguous()
    out["kat_bones"] = kat_len
    out["kat_pose_identity"] = dec(ident, kat_len, torch.zeros(1, 3))
    out["kat_t_pose"] = dec.build_t_pose_from_bone_lengths(kat_len)
    r4 = torch.randn(20, 17, 4, generator=g)
    dec4 = ref.PoseDecoder(ref.make_skeleton(), rot_rep_dim=4)
    out["rot4d"] = r4
    out["poses_4d"] = dec4(r4, bones, torch.zeros(20, 3))
    torch.save(out, os.path.join(OUT, "decoder.pt"))


def loss_golden(ref):
    M = ref.metrics
    g = torch.Generator().manual_seed(99)
    b, k, t = 3, 5, 11
    y = 0.3 * torch.randn(b, t, 17, 3, generator=g)
    y[:, :, 0] = 0
    hyp = (y[:, None] + 0.1 * torch.randn(b, k, t, 17, 3, generator=g)).requires_grad_(True)
    logits = torch.randn(b, k, t, 1, generator=g, requires_grad=True)
    scores = logits.softmax(dim=1)
    w = M.STANDARD_H36M_WEIGHTS
    out = {"hyp": hyp.detach().clone(), "logits": logits.detach().clone(), "scores": scores.detach().clone(), "y": y}
    for name, weights, squared in (("w", w, False), ("u", None, False), ("wsq", w, True)):
        v, i = M.wta_l2_loss_and_activate_head(hyp, y, weights, squared)
        out[f"wta_val_{name}"], out[f"wta_idx_{name}"] = v.detach().clone(), i.clone()
        tot, bce = M.wta_with_scoring_loss(hyp, scores, y, 0.1, weights, squared)
        out[f"score_total_{name}"], out[f"score_bce_{name}"] = tot.detach().clone(), bce.detach().clone()
    out["vel"] = M.mean_velocity_error(hyp, y, axis=2).detach().clone()
    out["vel_sq"] = M.mean_velocity_error(hyp, y, axis=2, squared=True).detach().clone()
    out["smooth_w"] = M.smoothness_regularization(hyp, w, axis=2).detach().clone()
    # the training loss of hpe/main_h36m_lifting.py:101-209 with config.yaml defaults, and its gradients
    wl = M.wta_l2_loss_and_activate_head(hypothesis=hyp, y=y, weights=w, squared=False)[0].mean()
    sr = M.wta_with_scoring_loss(hypothesis=hyp, scores=scores, y=y, beta=0.1, weights=w, squared=False)[1]
    vl = 2.0 * M.mean_velocity_error(predicted=hyp, target=y, squared=False, axis=2)
    sg = 0.5 * M.smoothness_regularization(prediction=hyp, weights=w, axis=2)
    loss = torch.zeros(1)
    for term in (wl, sr, vl, sg):
        loss = loss + term
    loss.backward()
    out.update(train_total=loss.detach().clone(), train_wloss=wl.detach().clone(), train_score_reg=sr.detach().clone(),
               train_vloss=vl.detach().clone(), train_sreg=sg.detach().clone(),
               grad_hyp=hyp.grad.clone(), grad_logits=logits.grad.clone())
    torch.manual_seed(0)
    m = ref.architectures.RMCLManifoldMixSTE(ref.make_skeleton(), num_frame=9, n_hyp=k)
    hd, sc = hyp.detach(), scores.detach()
    out["agg_weighted"] = m.aggregate(hd, sc, "weighted_ave")
    out["agg_best"] = m.aggregate(hd, sc, "best_score")
    out["agg_oracle_val"], out["agg_oracle_pose"] = m.aggregate(hd, mode="oracle", ground_truth=y)
    out["mpjpe_sum"] = M.mpjpe_error(out["agg_weighted"], y, "sum")
    out["mpjpe_avg"] = M.mpjpe_error(out["agg_weighted"], y, "average")
    torch.save(out, os.path.join(OUT, "loss.pt"))


def forward_golden(ref):
    out = {}
    for tag, T, k, perturb in (("t27k5_init", 27, 5, False), ("t27k5_pert", 27, 5, True), ("t9k1_pert", 9, 1, True)):
        torch.manual_seed(42)
        m = ref.architectures.RMCLManifoldMixSTE(ref.make_skeleton(), num_frame=T, n_hyp=k, drop_path_rate=0.1).eval()
        if perturb:
            g = torch.Generator().manual_seed(7)
            with torch.no_grad():
                for p in m.parameters():
                    p.add_(torch.randn(p.shape, generator=g) * 0.02)
        x = 0.3 * torch.randn(1, T, 17, 2, generator=torch.Generator().manual_seed(1234))
        with torch.no_grad():
            poses, scores = m(x)
            rot, _ = m.rotations_module(x)
            bones = m.segments_module(x)
        out[tag] = {"T": T, "K": k, "perturb": perturb, "x": x, "poses": poses, "scores": scores,
                    "rot6d": rot, "bones": bones, "checksum": weights_checksum(m.state_dict()),
                    "keys": [(n, tuple(t.shape)) for n, t in m.state_dict().items()]}
    # seeded synthetic weights (oracle.make_state_dict) loaded INTO the reference model
    sys.path.insert(0, ROOT)
    from oracle.manipose_oracle import make_state_dict
    sd = make_state_dict(num_frame=27, n_hyp=5, seed=3)
    m = ref.architectures.RMCLManifoldMixSTE(ref.make_skeleton(), num_frame=27, n_hyp=5).eval()
    m.load_state_dict(sd)
    x = 0.3 * torch.randn(2, 27, 17, 2, generator=torch.Generator().manual_seed(77))
    with torch.no_grad():
        poses, scores = m(x)
    out["t27k5_synth"] = {"T": 27, "K": 5, "seed": 3, "x": x, "poses": poses, "scores": scores,
                          "checksum": weights_checksum(sd)}
    torch.save(out, os.path.join(OUT, "forward.pt"))


def consistency_golden(ref):
    """the reference's own bone-length / MPSCE / MPSSE functions on seeded poses."""
    sk = ref.make_skeleton()
    M = ref.metrics
    out = {}
    for tag, b, l in (("b3_l27", 3, 27), ("b1_l243", 1, 243), ("b2_l700", 2, 700)):
        poses = 0.3 * torch.randn(b, l, 17, 3, generator=torch.Generator().manual_seed(b * 1000 + l))
        jc = poses.permute(0, 3, 2, 1)
        e = {"poses": poses, "bone_len": M.measure_bones_length(jc, sk.bones)}
        for mode in ("average", "sum", "std", "min", "max"):
            e[f"stc_{mode}"] = M.segments_time_consistency(jc, sk, mode)
        for mode in ("average", "sum", "std"):
            e[f"stc_pb_{mode}"] = M.segments_time_consistency_per_bone(jc, sk, mode)
        for mode in ("average", "sum"):
            for squared in (True, False):
                e[f"sym_{mode}_{int(squared)}"] = M.sagittal_symmetry(jc, sk, mode, squared)
                e[f"sym_pb_{mode}_{int(squared)}"] = M.sagittal_symmetry_per_bone(jc, sk, mode, squared)
        out[tag] = e
    torch.save(out, os.path.join(OUT, "consistency.pt"))


def tta_golden(ref):
    """flip test-time augmentation exactly as hpe/eval_utils.py:51-142 composes the reference's own pieces."""
    sys.path.insert(0, ROOT)
    from oracle.manipose_oracle import make_state_dict
    from mh_so3_hpe.augmentations.functional import pose_flip
    sk = ref.make_skeleton()
    sd = make_state_dict(num_frame=27, n_hyp=5, seed=5)
    m = ref.architectures.RMCLManifoldMixSTE(sk, num_frame=27, n_hyp=5).eval()
    m.load_state_dict(sd)
    x = 0.3 * torch.randn(2, 27, 17, 2, generator=torch.Generator().manual_seed(31))
    out = {"T": 27, "K": 5, "seed": 5, "x": x.clone()}
    with torch.no_grad():
        for mode in ("weighted_ave", "best_score"):
            pred = m.aggregate(*m(x.clone()), mode=mode)
            x_f = pose_flip(poses_tuple=(x.clone(),), skeleton=sk)[0]
            pred_f = m.aggregate(*m(x_f), mode=mode)
            pred_f = pose_flip(poses_tuple=(pred_f,), skeleton=sk)[0]
            out[mode] = (pred + pred_f) / 2
    torch.save(out, os.path.join(OUT, "tta.pt"))


def procrustes_golden(ref):
    """the reference's numpy P-MPJPE on seeded pose pairs (noisy, similarity-transformed, mirrored)."""
    from mh_so3_hpe.metrics.mean_joint_errors import p_mpjpe
    g = torch.Generator().manual_seed(17)
    y = 0.3 * torch.randn(3, 27, 17, 3, generator=g)
    y[:, :, 0] = 0
    cases = {"noisy": (y + 0.05 * torch.randn(3, 27, 17, 3, generator=g), y)}
    q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g))
    cases["similarity"] = (1.7 * y @ q + torch.tensor([0.3, -0.2, 0.9]) + 0.01 * torch.randn(3, 27, 17, 3, generator=g), y)
    mirrored = y.clone()
    mirrored[..., 0] *= -1
    cases["mirrored"] = (mirrored + 0.02 * torch.randn(3, 27, 17, 3, generator=g), y)
    out = {name: {"pred": p, "target": t, "p_mpjpe": float(p_mpjpe(p, t))} for name, (p, t) in cases.items()}
    # 3DPCK / AUC (pck.py:77-198) on millimetre-scale points, including errors that sit exactly on thresholds
    from mh_so3_hpe.metrics.pck import keypoint_3d_pck, keypoint_3d_auc
    gt = 300.0 * torch.randn(400, 17, 3, generator=g)
    pred = gt + 60.0 * torch.randn(400, 17, 3, generator=g)
    pred[0, :, :] = gt[0, :, :]
    pred[0, :, 0] += torch.arange(17, dtype=torch.float32) * 10.0          # errors 0, 10, ..., 160: exact multiples of 5
    out["pck"] = {"pred": pred, "gt": gt, "pck150": float(keypoint_3d_pck(pred, gt, threshold=150)),
                  "pck50": float(keypoint_3d_pck(pred, gt, threshold=50)), "auc": float(keypoint_3d_auc(pred, gt))}
    torch.save(out, os.path.join(OUT, "procrustes.pt"))


def _load_reference_evaluate():
    """hpe/eval_utils.py imports omegaconf only for a type hint (:8): stub it, then load the module from the reference tree."""
    import importlib.util
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


def evaluate_golden(ref):
    """The whole of ``evaluate`` (hpe/eval_utils.py:16-203), unmodified, on synthetic weights loaded into the reference model."""
    import types
    from oracle import manipose_oracle as O
    ev = _load_reference_evaluate()
    sk = ref.make_skeleton()
    t, k, seed = 9, 3, 5
    m = ref.architectures.RMCLManifoldMixSTE(sk, num_frame=t, n_hyp=k, drop_path_rate=0.1).eval()
    m.load_state_dict(O.make_state_dict(num_frame=t, n_hyp=k, seed=seed))
    g = torch.Generator().manual_seed(8)
    batches = [(0.3 * torch.randn(b, t, 17, 2, generator=g), 0.3 * torch.randn(b, t, 17, 3, generator=g)) for b in (2, 3)]
    out = {"T": t, "K": k, "seed": seed, "batches": batches}
    for tta in (False, True):
        cfg = types.SimpleNamespace(train=types.SimpleNamespace(tta=tta))
        res = ev.evaluate(m, [(x.clone(), y.clone()) for x, y in batches], "cpu", cfg, sk, return_hyps=False, compute_oracle=True)
        out[f"tta{int(tta)}"] = {"predictions": [p.clone() for p in res[0]], "performance": float(res[2]), "oracle_mpjpe": float(res[3]),
                                 "psoracle_mpjpe": float(res[4]), "oracle_preds": [p.clone() for p in res[5]]}
    torch.save(out, os.path.join(OUT, "evaluate.pt"))


def windows_golden(ref):
    """Items of the reference ``PoseSequenceGenerator`` (hpe/mh_so3_hpe/data/generators.py:45-219) under fixed torch / numpy seeds: plain
    windows with the replicate-padded tail, random starts, every occlusion pattern, the noisy input and the PoseFlip transform."""
    import numpy as np
    from mh_so3_hpe.data.generators import PoseSequenceGenerator
    from mh_so3_hpe.augmentations.transforms import PoseFlip
    rng = np.random.default_rng(3)
    lens = [60, 45, 100, 7]
    p3 = [rng.standard_normal((n, 17, 3)) for n in lens]     # float64: the reference's in-place noise / flip cannot leak into the arrays
    p2 = [rng.standard_normal((n, 17, 2)) for n in lens]
    out = {"p3": p3, "p2": p2, "seq_len": 20, "cases": []}
    cases = [dict(drop_last=False, random_start=False, miss_type="no_miss", flip=None, order=list(range(12)))]
    for mt in ("random", "random_left_arm_right_leg", "structured_joint", "structured_frame", "noisy", "all"):
        cases.append(dict(drop_last=True, random_start=True, miss_type=mt, flip=None, order=[4, 0, 7, 2, 5, 1]))
    cases.append(dict(drop_last=True, random_start=True, miss_type="random", flip=0.5, order=[4, 0, 7, 2, 5, 1, 3, 6]))
    for i, c in enumerate(cases):
        tr = PoseFlip(skeleton=ref.make_skeleton(), probability=c["flip"]) if c["flip"] is not None else None
        seqs3 = p3[:3] if c["random_start"] else p3           # random starts need sequences longer than one window
        seqs2 = p2[:3] if c["random_start"] else p2
        gen = PoseSequenceGenerator(seqs3, seqs2, None, seq_len=20, random_start=c["random_start"], drop_last=c["drop_last"],
                                    miss_type=c["miss_type"], miss_rate=0.3, noise_sigma=0.05, transform=tr)
        torch.manual_seed(100 + i)
        np.random.seed(100 + i)
        items = [gen[j] for j in c["order"]]
        out["cases"].append({**c, "n_seqs": len(seqs3), "seed": 100 + i, "length": len(gen),
                             "items": [(torch.as_tensor(a).clone(), torch.as_tensor(b).clone()) for a, b in items]})
    torch.save(out, os.path.join(OUT, "windows.pt"))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    ref = load_reference()
    makers = {"decoder": decoder_golden, "loss": loss_golden, "forward": forward_golden, "consistency": consistency_golden, "tta": tta_golden,
              "procrustes": procrustes_golden, "evaluate": evaluate_golden, "windows": windows_golden}
    for name in (sys.argv[1:] or list(makers)):          # python scripts/make_goldens.py [fixture ...]
        makers[name](ref)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
