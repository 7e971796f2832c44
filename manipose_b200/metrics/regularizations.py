"""Regularizers / pose-consistency metrics with the reference's signatures (hpe/mh_so3_hpe/metrics/regularizations.py,
metrics/utils.py): smoothness on the loss reduction kernel, bone-length statistics (MPSCE, MPSSE) on mp_pose_consistency."""
import torch

from .. import _lib as L
from .. import ops
from ..data.skeleton import skeleton_tables
from .losses import _as_hyp


def smoothness_regularization(prediction: torch.Tensor, weights: torch.Tensor = None, axis: int = 1) -> torch.Tensor:
    """mean over everything of w_j * (d/dt prediction)^2."""
    if prediction.dim() == 5 and axis != 2 or prediction.dim() == 4 and axis != 1:
        raise NotImplementedError(f"smoothness_regularization(axis={axis}) on rank {prediction.dim()} is not built")
    if weights is not None:
        assert weights.shape[0] == prediction.shape[-2]
    hyp = prediction if prediction.dim() == 5 else prediction.unsqueeze(1)
    # the term does not involve a target: pass a zero target of the right shape (read once, 204 B/frame)
    y = torch.zeros((hyp.shape[0], hyp.shape[2], hyp.shape[3], 3), dtype=torch.float32, device=hyp.device)
    return ops.loss_terms(hyp, None, y, weights, False)[0][L.MP_TERM_SMOOTH]


# ------------------------------------------------------------------------------------------------ pose consistency (MPSCE / MPSSE)
def _poses_of(joints_coords: torch.Tensor, skeleton) -> torch.Tensor:
    """The reference passes ``poses.permute(0, 3, 2, 1)`` = [B, 3, J, L]; the kernel reads [B, L, J, 3] (a view of the same memory when
    the caller permuted a contiguous pose tensor, a copy otherwise)."""
    if joints_coords.dim() != 4 or joints_coords.shape[1] != 3:
        raise AssertionError("joints_coords must be [batch, 3, num_joints, series_length]")
    assert joints_coords.shape[2] == len(skeleton.bones) + 1
    if torch.is_grad_enabled() and joints_coords.requires_grad:
        raise NotImplementedError("pose-consistency metrics are evaluation-only here (the rigid_seg_reg training term is off in every "
                                  "BASELINE config)")
    ops.set_skeleton(*skeleton_tables(skeleton))
    return joints_coords.permute(0, 3, 2, 1)


def measure_bones_length(joints_coords: torch.Tensor, skeleton_bones) -> torch.Tensor:
    """metrics/utils.py:4-20 -> [B, num_bones, L].  ``skeleton_bones`` must be the (joint, parent) list of the 17-joint tree."""
    want = tuple((j + 1, ops_parent(j + 1)) for j in range(ops.BONES))
    if tuple((int(j), int(p)) for j, p in skeleton_bones) != want:
        raise NotImplementedError("measure_bones_length is built for the H36M-17 / MPI-INF-3DHP tree (SURVEY.md §A.1)")
    if joints_coords.dim() != 4 or joints_coords.shape[1] != 3 or joints_coords.shape[2] != ops.J:
        raise AssertionError("joints_coords must be [batch, 3, 17, series_length]")
    return ops.pose_consistency(joints_coords.permute(0, 3, 2, 1), with_bone_lengths=True)[4]


def ops_parent(j: int) -> int:
    return (-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 9, 8, 11, 12, 8, 14, 15)[j]


def _segments_stat(joints_coords, skeleton, mode):
    if mode not in ("average", "sum", "std", "min", "max"):
        raise ValueError(f"Unexpected value for 'mode' encoutered: {mode}.Accepted values are 'average', 'sum' and 'std.")
    _, seg_var, _, _, _ = ops.pose_consistency(_poses_of(joints_coords, skeleton))
    stat = seg_var.sqrt() if mode == "std" else seg_var
    agg = {"average": torch.mean, "std": torch.mean, "sum": torch.sum, "min": torch.min, "max": torch.max}[mode]
    return stat, agg


def segments_time_consistency(joints_coords: torch.Tensor, skeleton, mode: str):
    """regularizations.py:38-48: aggregate over (batch, bone) of the unbiased variance (std for mode='std') over time of the bone lengths."""
    stat, agg = _segments_stat(joints_coords, skeleton, mode)
    return agg(stat)


def segments_time_consistency_per_bone(joints_coords: torch.Tensor, skeleton, mode: str):
    """regularizations.py:51-61: the same, aggregated over the batch only."""
    stat, agg = _segments_stat(joints_coords, skeleton, mode)
    return agg(stat, dim=0)


def _sagittal(joints_coords, skeleton, mode, squared):
    if mode not in ("average", "sum"):
        raise ValueError(f"Unexpected value for 'mode' encoutered: {mode}.Accepted values are 'average' and 'sum'.")
    if tuple(skeleton.bones_left) != (3, 4, 5, 10, 11, 12) or tuple(skeleton.bones_right) != (0, 1, 2, 13, 14, 15):
        raise NotImplementedError("sagittal_symmetry is built for the H36M-17 / MPI-INF-3DHP left/right bone pairs")
    poses = _poses_of(joints_coords, skeleton)
    _, _, sym_abs, sym_sq, _ = ops.pose_consistency(poses)
    per_pair = sym_sq if squared else sym_abs           # [B, 6]: means over time
    return per_pair, poses.shape[1]


def sagittal_symmetry(joints_coords: torch.Tensor, skeleton, mode: str, squared: bool = True):
    """regularizations.py:130-140: mean / sum over (batch, pair, time) of |len[left] - len[right]| (squared by default)."""
    per_pair, n_frames = _sagittal(joints_coords, skeleton, mode, squared)
    return per_pair.mean() if mode == "average" else per_pair.sum() * n_frames


def sagittal_symmetry_per_bone(joints_coords: torch.Tensor, skeleton, mode: str, squared: bool = True):
    """regularizations.py:143-157: per left/right pair, aggregated over batch and time."""
    per_pair, n_frames = _sagittal(joints_coords, skeleton, mode, squared)
    return per_pair.mean(dim=0) if mode == "average" else per_pair.sum(dim=0) * n_frames
