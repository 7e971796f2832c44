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
