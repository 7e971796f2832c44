"""Evaluation epilogue of the lifting path: flip test-time augmentation + hypothesis aggregation,
hpe/eval_utils.py:51-142 for ``RMCLManifoldMixSTE``.

The reference runs the model twice (input, flipped input), aggregates each, flips the second prediction back with
``pose_flip`` and averages: two forwards and ~30 small launches with host-built index tensors.  Here both inputs go through ONE
forward (clips are independent, so stacking them changes nothing per clip) and one kernel (``mp_aggregate_tta``) does both
aggregations, the un-flip and the average."""
import torch

from . import _lib as L
from . import ops

_MODES = {"weighted_ave": L.MP_AGG_WEIGHTED_AVE, "best_score": L.MP_AGG_BEST_SCORE}


def flip_input(x: torch.Tensor, skeleton) -> torch.Tensor:
    """``pose_flip`` of augmentations/functional.py:7-28 as a pure function: negate the horizontal coordinate, swap left / right joints."""
    perm = list(range(skeleton.num_joints))
    for l, r in zip(skeleton.joints_left, skeleton.joints_right):
        perm[l], perm[r] = r, l
    out = x[..., perm, :].clone()
    out[..., 0] *= -1
    return out


@torch.no_grad()
def lift_with_tta(model, x: torch.Tensor, mode: str = "weighted_ave") -> torch.Tensor:
    """x [B,L,17,2] -> aggregated 3-D poses [B,L,17,3] with flip test-time augmentation (``config.train.tta``)."""
    if mode not in _MODES:
        raise ValueError(f"Only best_score and weighted_ave modes are implemented.Got {mode}.")
    ops._need_cuda(x)
    skeleton = model.decoder.skeleton
    if tuple(skeleton.joints_left) != (4, 5, 6, 11, 12, 13) or tuple(skeleton.joints_right) != (1, 2, 3, 14, 15, 16):
        raise NotImplementedError("the TTA epilogue is built for the H36M-17 / MPI-INF-3DHP left / right joints")
    both = torch.cat([x, flip_input(x, skeleton)], dim=0)
    was_training = model.training
    model.eval()
    try:
        poses, scores = model(both)
    finally:
        model.train(was_training)
    return ops.aggregate_tta(poses, scores.reshape(scores.shape[:3]), _MODES[mode])


def _flip_poses(p: torch.Tensor, skeleton) -> torch.Tensor:
    return flip_input(p, skeleton)


@torch.no_grad()
def evaluate(model, loader, device, config, skeleton, return_hyps: bool = False, compute_oracle: bool = True):
    """Drop-in for ``evaluate`` of hpe/eval_utils.py:16-203 with an ``RMCLManifoldMixSTE`` from this package: same arguments, same return
    tuple — ``(all_predictions, all_target, performance)`` or, with ``compute_oracle``, additionally ``(oracle_mpjpe, psoracle_mpjpe,
    all_oracle_preds)`` — and the same normalisations (the non-TTA oracle figure is divided by J once more than the others, :57-64,186-197).

    Differences in HOW: with ``config.train.tta`` the input and its flip go through ONE forward; every aggregation / MPJPE is a kernel of
    this library (``mp_aggregate``, ``mp_mpjpe``); the per-batch scalars stay on the device until the end of the loop (the reference
    synchronises with ``.item()`` / ``.cpu()`` every batch, :172-173), so batches pipeline."""
    from .architectures.rmcl_manifold_mix_ste import RMCLManifoldMixSTE
    if not isinstance(model, RMCLManifoldMixSTE):
        raise TypeError("manipose_b200.evaluation.evaluate is built for RMCLManifoldMixSTE (the lifting hot path)")
    tta = bool(config.train.tta)
    was_training = model.training
    model.eval()
    all_pred, all_target, all_oracle = [], [], []
    sums, means, oracle_sums, ps_sums = [], [], [], []
    n, l, j = 0, 0, skeleton.num_joints
    try:
        for input_2d, target_3d in loader:
            b, l, j, _ = target_3d.shape
            y = target_3d.to(device).float().contiguous()
            x = input_2d.to(device).float().contiguous()
            if tta:
                poses2, scores2 = model(torch.cat([x, flip_input(x, skeleton)], dim=0))
                poses, scores, poses_f, scores_f = poses2[:b], scores2[:b], poses2[b:], scores2[b:]
                pred = ops.aggregate_tta(poses2, scores2.reshape(scores2.shape[:3]), L.MP_AGG_WEIGHTED_AVE)
            else:
                poses, scores = model(x)
                pred = model.aggregate(poses, scores)
            if compute_oracle:
                val, oracle_preds = model.aggregate(poses, mode="oracle", ground_truth=y)
                ps_preds = model.aggregate(poses, scores, mode="best_score")
                if tta:
                    hyp_f = _flip_poses(poses_f, skeleton)
                    _, oracle_f = model.aggregate(hyp_f, mode="oracle", ground_truth=y)
                    oracle_preds = (oracle_preds + oracle_f) / 2
                    oracle_sums.append(ops.mpjpe(oracle_preds, y)[0] / j)
                    ps_tta = (ps_preds + model.aggregate(hyp_f, scores_f, mode="best_score")) / 2
                    ps_sums.append(ops.mpjpe(ps_tta, y)[0] / j)
                else:
                    oracle_sums.append(val.sum() / j)
                    ps_sums.append(ops.mpjpe(ps_preds, y)[0] / j)
                all_oracle.append(oracle_preds * 1000)
            n += b
            if return_hyps:
                hyp = model.concat_hyp_and_scores(poses, scores)
                hyp[..., :-1] *= 1000
                all_pred.append(hyp)
            else:
                all_pred.append(pred * 1000)
            err = ops.mpjpe(pred, y)                                      # (sum, mean) in metres
            sums.append(err[0])
            means.append(err[1])
            all_target.append(y)
    finally:
        model.train(was_training)
    batch_no = len(sums)
    print("Average MPJPE:", float(torch.stack(means).sum() * 1000) / batch_no if batch_no else float("nan"))
    performance = (torch.stack(sums).double().sum() * 1000 / (n * l * j)).cpu().numpy()
    if not compute_oracle:
        return all_pred, all_target, performance
    oracle_total = torch.stack(oracle_sums).sum() / (n * l) * 1000
    ps_total = torch.stack(ps_sums).sum() / (n * l) * 1000
    return all_pred, all_target, performance, oracle_total, ps_total, all_oracle


def stack_lifted(predictions) -> "np.ndarray":
    """Post-processing of ``lift_action`` (hpe/eval_utils.py:241-251): the per-batch predictions of ``evaluate`` (millimetres) -> one numpy
    array in metres, frames flattened: [N*L, J, 3] for aggregated poses, [N*L, K, J, 4] (pose + score) for ``return_hyps``."""
    import numpy as np
    out = torch.cat(list(predictions), dim=0).detach().cpu().numpy()
    if out.ndim == 4:
        n, l, j, _ = out.shape
        return out.reshape(n * l, j, 3) / 1000
    out = np.transpose(out, (0, 2, 1, 3, 4))
    n, l, _, j, _ = out.shape
    out = out.reshape(n * l, -1, j, 4)
    out[..., :-1] /= 1000
    return out


def lift_action(data_loader, model, device, config, skeleton, return_hyps):
    """Drop-in for ``lift_action`` of hpe/eval_utils.py:226-251 (what hpe/viz.py calls): ``evaluate`` + ``stack_lifted``."""
    return stack_lifted(evaluate(model=model, loader=data_loader, device=device, config=config, skeleton=skeleton, return_hyps=return_hyps)[0])
