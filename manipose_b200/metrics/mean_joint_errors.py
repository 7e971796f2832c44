"""mpjpe_error (hpe/mh_so3_hpe/metrics/mean_joint_errors.py:8-36) on the device reduction kernel."""
import torch

from .. import ops


def mpjpe_error(batch_imp: torch.Tensor, batch_gt: torch.Tensor, mode: str):
    assert batch_imp.shape[-1] == batch_gt.shape[-1] == 3
    if mode == "average":
        return ops.mpjpe(batch_imp, batch_gt)[1]
    if mode == "sum":
        return ops.mpjpe(batch_imp, batch_gt)[0]
    if mode == "no_agg":
        raise NotImplementedError("mpjpe_error(mode='no_agg') is only used by offline per-action analytics (out of scope, SURVEY.md §2)")
    raise ValueError(f"Unexpected value for 'mode' encoutered: {mode}.Accepted values are 'average' and 'sum'.")


def p_mpjpe(predicted: torch.Tensor, target: torch.Tensor):
    """mean_joint_errors.py:144-189: MPJPE after rigid alignment (scale, rotation, translation), "Protocol #2".  The reference copies
    both tensors to the host for numpy's batched SVD; this one solves every frame's 3 x 3 Procrustes problem on the device and returns
    a Python float like the reference's ``np.mean``."""
    assert predicted.shape == target.shape
    assert predicted.shape[-1] == target.shape[-1] == 3
    return float(ops.p_mpjpe(predicted, target)[1])
