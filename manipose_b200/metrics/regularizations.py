"""smoothness_regularization (hpe/mh_so3_hpe/metrics/regularizations.py:160-174) on the loss reduction kernel."""
import torch

from .. import _lib as L
from .. import ops
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
