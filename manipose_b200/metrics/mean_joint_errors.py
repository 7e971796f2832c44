"""Joint-error metrics with the reference's signatures (hpe/mh_so3_hpe/metrics/mean_joint_errors.py:8-189) on the device
reduction kernels (mp_mpjpe, mp_point_errors, mp_p_mpjpe, mp_pose_consistency).  CPU tensors and numpy arrays are moved to the
device (the analytics of the drivers call these on whatever they hold); there is no CPU arithmetic."""
import torch

from .. import _lib as L
from .. import ops


def _dev(t):
    """Inputs of the analytics functions: device tensors pass through, CPU tensors / numpy arrays go to the current device."""
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t)
    return t if t.is_cuda else t.cuda()


def _mode_check(mode):
    if mode not in ("average", "sum", "no_agg"):
        raise ValueError(f"Unexpected value for 'mode' encoutered: {mode}.Accepted values are 'average' and 'sum'.")


def _pair(batch_imp, batch_gt):
    batch_imp, batch_gt = _dev(batch_imp), _dev(batch_gt)
    assert batch_imp.shape[-1] == batch_gt.shape[-1] == 3
    return batch_imp, batch_gt


def mpjpe_error(batch_imp, batch_gt, mode: str):
    """:31-36  torch.norm(gt - imp, 2, 1) over all 3-D points -> mean | sum | the per-point vector."""
    batch_imp, batch_gt = _pair(batch_imp, batch_gt)
    _mode_check(mode)
    if mode == "average":
        return ops.mpjpe(batch_imp, batch_gt)[1]
    if mode == "sum":
        return ops.mpjpe(batch_imp, batch_gt)[0]
    return ops.point_errors(batch_imp, batch_gt, L.MP_ERR_L2, per_elem=True)[0]


def mse_error(batch_imp, batch_gt, mode: str):
    """:39-44  squared distance of every 3-D point -> mean | sum | the per-point vector."""
    batch_imp, batch_gt = _pair(batch_imp, batch_gt)
    _mode_check(mode)
    n = batch_imp.numel() // 3
    if mode == "no_agg":
        return ops.point_errors(batch_imp, batch_gt, L.MP_ERR_SQ, per_elem=True)[0]
    return ops.point_errors(batch_imp, batch_gt, L.MP_ERR_SQ, cols=1, scale=1.0 / n if mode == "average" else 1.0)[1][0]


def _jointwise(batch_imp, batch_gt, mode, err_mode):
    batch_imp, batch_gt = _pair(batch_imp, batch_gt)
    _mode_check(mode)
    j = batch_gt.shape[-2]
    rows = batch_gt.numel() // (3 * j)
    if mode == "no_agg":
        return ops.point_errors(batch_imp, batch_gt, err_mode, per_elem=True)[0].view(rows, j)
    return ops.point_errors(batch_imp, batch_gt, err_mode, cols=j, scale=1.0 / rows if mode == "average" else 1.0)[1]


def jointwise_error(batch_imp, batch_gt, mode: str):
    """:47-62  per-joint L2 error aggregated over every pose -> [J]."""
    return _jointwise(batch_imp, batch_gt, mode, L.MP_ERR_L2)


def jointwise_mse(batch_imp, batch_gt, mode: str):
    """:65-80  per-joint squared error aggregated over every pose -> [J]."""
    return _jointwise(batch_imp, batch_gt, mode, L.MP_ERR_SQ)


def coordwise_error(batch_imp, batch_gt, mode: str):
    """:132-141  |gt - imp| per coordinate aggregated over every point -> [3]."""
    batch_imp, batch_gt = _pair(batch_imp, batch_gt)
    _mode_check(mode)
    n = batch_imp.numel() // 3
    if mode == "no_agg":
        return ops.point_errors(batch_imp, batch_gt, L.MP_ERR_ABS, per_elem=True)[0].view(n, 3)
    return ops.point_errors(batch_imp, batch_gt, L.MP_ERR_ABS, cols=3, scale=1.0 / n if mode == "average" else 1.0)[1]


def segments_len_err(batch_imp, batch_gt, skeleton, mode: str, signed: bool = True):
    """:83-129  gt - predicted bone lengths ([B, 3, J, L] inputs like measure_bones_length) -> mean | sum over (frame, bone), or
    the [B*L, num_bones] matrix for 'no_agg'."""
    from .regularizations import measure_bones_length
    batch_imp, batch_gt = _dev(batch_imp), _dev(batch_gt)
    _mode_check(mode)
    pred_len = measure_bones_length(batch_imp, skeleton.bones)       # [B, bones, L]
    gt_len = measure_bones_length(batch_gt, skeleton.bones)
    err_mode = L.MP_ERR_DIFF if signed else L.MP_ERR_ABS
    b, nb, l = pred_len.shape
    if mode == "no_agg":
        diff = ops.point_errors(pred_len, gt_len, err_mode, per_elem=True)[0]
        return diff.view(b, nb, l).permute(0, 2, 1).reshape(b * l, nb)
    n = pred_len.numel()
    return ops.point_errors(pred_len, gt_len, err_mode, cols=1, scale=1.0 / n if mode == "average" else 1.0)[1][0]


def p_mpjpe(predicted, target):
    """:144-189: MPJPE after rigid alignment (scale, rotation, translation), "Protocol #2".  The reference copies both tensors to the
    host for numpy's batched SVD; this one solves every frame's 3 x 3 Procrustes problem on the device and returns a Python float
    like the reference's ``np.mean``."""
    predicted, target = _dev(predicted), _dev(target)
    assert predicted.shape == target.shape
    assert predicted.shape[-1] == target.shape[-1] == 3
    return float(ops.p_mpjpe(predicted, target)[1])
