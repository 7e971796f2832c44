"""3DPCK / AUC with the reference's signatures (hpe/mh_so3_hpe/metrics/pck.py:77-198) on the device histogram kernel.

Only what the drivers use is built (hpe/main_3dhp.py:882-910): ``alignment='none'`` and ``mask=None``; other options raise."""
import torch

from .. import ops


def _check(pred, gt, mask, alignment):
    if mask is not None:
        raise NotImplementedError("keypoint_3d_pck / keypoint_3d_auc with a visibility mask are not built (the drivers pass mask=None)")
    if alignment != "none":
        if alignment in ("procrustes", "scale"):
            raise NotImplementedError(f"alignment='{alignment}' is not built (the drivers use 'none'; see metrics.p_mpjpe for Procrustes)")
        raise ValueError(f"Invalid value for alignment: {alignment}")


def _dev(t):
    """numpy arrays / CPU tensors (the reference's handle_tensors converts the other way, pck.py:83-89) go to the device."""
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t)
    return t if t.is_cuda else t.cuda()


def keypoint_3d_pck(pred, gt, mask=None, alignment="none", threshold=150.0):
    """pck.py:77-141: percentage of joints whose error is below ``threshold`` (default 150 mm)."""
    _check(pred, gt, mask, alignment)
    return float(ops.pck_auc(_dev(pred), _dev(gt), threshold)[0])


def keypoint_3d_auc(pred, gt, mask=None, alignment="none"):
    """pck.py:144-198: area under the PCK curve for thresholds linspace(0, 150, 31)."""
    _check(pred, gt, mask, alignment)
    return float(ops.pck_auc(_dev(pred), _dev(gt))[1])
