"""Loss functions of the hot path with the reference's signatures (hpe/mh_so3_hpe/metrics/losses.py), computed by the
warp-tile reduction kernels of libmanipose_sm100 (mp_loss_fwd / mp_loss_bwd / mp_wta_fwd) with matching backward.

Supported layouts are the ones the drivers use (hpe/main_h36m_lifting.py:101-209): hypothesis [B,H,L,17,3] with target
[B,L,17,3] (time on axis 2), and single predictions [B,L,17,3] with time on axis 1 (treated as H = 1).  Anything else
raises — there is no PyTorch fallback.
"""
from typing import List, Optional, Tuple

import torch

from .. import _lib as L
from .. import ops

# From MixSTE code (losses.py:6-11)
STANDARD_H36M_WEIGHTS = torch.Tensor([1, 1, 2.5, 2.5, 1, 2.5, 2.5, 1, 1, 1, 1.5, 1.5, 4, 4, 1.5, 4, 4])
STANDARD_HEVA_WEIGHTS = torch.Tensor([1, 1, 2.5, 2.5, 1, 2.5, 2.5, 1, 1.5, 1.5, 4, 4, 1.5, 4, 4])


def _as_hyp(prediction: torch.Tensor, target: torch.Tensor, axis: Optional[int] = None):
    """-> (hyp [B,H,L,17,3], y [B,L,17,3]) views for the layouts the kernels handle."""
    if prediction.dim() == 5:
        if target.dim() == 5:   # reference callers pass y[:, None].expand_as(hypothesis)
            target = target[:, 0]
        if axis is not None and axis != 2:
            raise NotImplementedError(f"time axis {axis} for a 5-D prediction is not built (drivers use axis=2)")
        return prediction, target
    if prediction.dim() == 4:
        if axis is not None and axis != 1:
            raise NotImplementedError(f"time axis {axis} for a 4-D prediction is not built (drivers use axis=1)")
        return prediction.unsqueeze(1), target
    raise NotImplementedError(f"prediction of rank {prediction.dim()} is not built (expected [B,H,L,J,3] or [B,L,J,3])")


def weighted_mpjpe_loss(prediction: torch.Tensor, target: torch.Tensor, weights: torch.Tensor = None,
                        dims: Optional[List[int]] = None) -> torch.Tensor:
    """losses.py:14-43: mean_j w_j ||pred - target||_2 (mean over everything when dims is None)."""
    if weights is not None:
        assert weights.shape[0] == target.shape[-2]
    hyp, y = _as_hyp(prediction, target)
    if dims is None:
        if hyp.shape[1] != 1:
            if torch.is_grad_enabled() and prediction.requires_grad:
                # every hypothesis contributes (no winner): fold H into the batch, so that each (clip, hypothesis) is a single-
                # hypothesis problem against its clip's target; the mean over (B*H, L) is the mean over (B, H, L)
                b, h = hyp.shape[:2]
                yy = y.unsqueeze(1).expand(b, h, *y.shape[1:]).reshape(b * h, *y.shape[1:])
                return ops.loss_terms(hyp.reshape(b * h, 1, *hyp.shape[2:]), None, yy, weights, False)[0][L.MP_TERM_WTA]
            # mean over B,H,L of the per-hypothesis error == mean of per_hyp
            _, _, per_hyp = ops.wta_fwd(hyp, y, weights, False, per_hyp=True)
            return per_hyp.mean()
        return ops.loss_terms(hyp, None, y, weights, False)[0][L.MP_TERM_WTA]
    if list(dims) == [3] and prediction.dim() == 5:
        if torch.is_grad_enabled() and prediction.requires_grad:
            raise NotImplementedError("per-hypothesis errors with gradients are exposed through wta_l2_loss_and_activate_head")
        return ops.wta_fwd(hyp, y, weights, False, per_hyp=True)[2]
    raise NotImplementedError(f"weighted_mpjpe_loss(dims={dims}) is not built")


def weighted_mse_loss(prediction: torch.Tensor, target: torch.Tensor, weights: torch.Tensor = None,
                      dims: Optional[List[int]] = None) -> torch.Tensor:
    """losses.py:46-72 (squared distances)."""
    if weights is None:
        if dims is not None:
            raise NotImplementedError("weighted_mse_loss(weights=None, dims=...) is F.mse_loss over everything in the reference (dims ignored)")
        # F.mse_loss(prediction, target) (losses.py:57-58) = the mean over every coordinate = the weighted form with unit weights
        weights = torch.ones(target.shape[-2])
    assert weights.shape[0] == target.shape[-2]
    hyp, y = _as_hyp(prediction, target)
    if dims is None and hyp.shape[1] == 1:
        return ops.loss_terms(hyp, None, y, weights, True)[0][L.MP_TERM_WTA]
    if dims is None:
        b, h = hyp.shape[:2]
        yy = y.unsqueeze(1).expand(b, h, *y.shape[1:]).reshape(b * h, *y.shape[1:])
        return ops.loss_terms(hyp.reshape(b * h, 1, *hyp.shape[2:]), None, yy, weights, True)[0][L.MP_TERM_WTA]
    if dims is not None and list(dims) == [4, 3] and prediction.dim() == 5 and not (torch.is_grad_enabled() and prediction.requires_grad):
        return ops.wta_fwd(hyp, y, weights, True, per_hyp=True)[2]
    raise NotImplementedError(f"weighted_mse_loss(dims={dims}) on shape {tuple(prediction.shape)} is not built")


def mean_velocity_error(predicted: torch.Tensor, target: torch.Tensor, axis: int = 1, squared: bool = False) -> torch.Tensor:
    """losses.py:75-101: mean ||d/dt pred - d/dt target||_2 (target broadcast over hypotheses)."""
    if predicted.dim() == target.dim():
        assert predicted.shape == target.shape
    hyp, y = _as_hyp(predicted, target, axis)
    w = torch.ones(ops.J) if squared else None   # the squared frame-error path needs weights; they do not enter this term
    return ops.loss_terms(hyp, None, y, w, squared)[0][L.MP_TERM_VEL]


def _l2_loss_per_hyp(hypothesis: torch.Tensor, y: torch.Tensor, weights: torch.Tensor = None, squared: bool = False) -> torch.Tensor:
    """losses.py:104-123 -> [B,H,L] (no gradient; the differentiable entry is wta_l2_loss_and_activate_head)."""
    return ops.wta_fwd(hypothesis, y, weights, squared, per_hyp=True)[2]


def wta_l2_loss_and_activate_head(hypothesis: torch.Tensor, y: torch.Tensor, weights: torch.Tensor = None,
                                  squared: bool = False) -> Tuple[torch.Tensor]:
    """losses.py:126-138: torch.min over hypotheses -> (values [B,L], int64 indices [B,L]); lowest index on ties."""
    if torch.is_grad_enabled() and hypothesis.requires_grad:
        _, val, idx = ops.loss_terms(hypothesis, None, y, weights, squared)
        return torch.return_types.min((val, idx))
    val, idx = ops.wta_fwd(hypothesis, y, weights, squared)
    return torch.return_types.min((val, idx))


def wta_with_scoring_loss(hypothesis: torch.Tensor, scores: torch.Tensor, y: torch.Tensor, beta: float,
                          weights: torch.Tensor = None, squared: bool = False):
    """losses.py:141-170: (wta.mean() + beta * BCE(scores, one_hot(winner)), beta * BCE); a bare scalar when beta == 0."""
    b, h, l = hypothesis.shape[:3]
    if beta == 0:
        return ops.loss_terms(hypothesis, None, y, weights, squared)[0][L.MP_TERM_WTA]
    terms = ops.loss_terms(hypothesis, scores.reshape(b, h, l), y, weights, squared)[0]
    scoring = beta * terms[L.MP_TERM_BCE]
    return terms[L.MP_TERM_WTA] + scoring, scoring


def training_loss(poses: torch.Tensor, scores: torch.Tensor, y: torch.Tensor, beta: float = 0.1, vel_w: float = 2.0,
                  smooth_w: float = 0.5, weights: torch.Tensor = STANDARD_H36M_WEIGHTS, squared: bool = False):
    """The whole objective of make_loss / compute_and_acc_loss (hpe/main_h36m_lifting.py:101-209, config.yaml:33-37) in ONE
    forward and ONE backward launch: returns (total, terms[8]) with terms = [wta, bce, velocity, smoothness, total, ...]."""
    b, h, l = poses.shape[:3]
    sc = scores.reshape(b, h, l) if scores is not None else None
    with ops.nvtx("manipose.loss"):
        terms = ops.loss_terms(poses, sc, y, weights, squared, beta, vel_w, smooth_w)[0]
    return terms[L.MP_TERM_TOTAL], terms
