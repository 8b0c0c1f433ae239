"""STEP loss: masked MAE + graph-structure BCE (reference: step/step_loss/step_loss.py:5-16,
basicts/metrics/mae.py:5-28)."""
import numpy as np
import torch


def masked_mae(preds: torch.Tensor, labels: torch.Tensor, null_val: float = np.nan) -> torch.Tensor:
    if np.isnan(null_val):
        mask = ~torch.isnan(labels)
    else:
        mask = (labels - null_val).abs() > 5e-5
    mask = mask.float()
    mask = mask / torch.mean(mask)
    mask = torch.where(torch.isnan(mask), torch.zeros_like(mask), mask)
    loss = torch.abs(preds - labels) * mask
    loss = torch.where(torch.isnan(loss), torch.zeros_like(loss), loss)
    return torch.mean(loss)


def step_loss(prediction, real_value, theta, priori_adj, gsl_coefficient, null_val=np.nan):
    """Reference signature (step/step_loss/step_loss.py:5).  On CUDA with the batch-invariant theta our STEP module
    returns (a stride-0 expanded [N,N] tensor), value and gradients come from the fused kernels
    (step_loss_fwd_bwd); any other input takes the elementwise formulation below (same math)."""
    if prediction.is_cuda and theta.dim() == 3 and theta.stride(0) == 0 and prediction.dtype == torch.float32:
        from step_b200 import ops
        return ops.FusedStepLoss.apply(prediction, real_value, theta[0], priori_adj, float(gsl_coefficient), float(null_val), 0.0, 1.0)
    # any other call shape: nn.BCELoss semantics (log clamped at -100, finite gradients at saturated theta) through the
    # library op, as the reference does (step_loss.py:10-13)
    bce = torch.nn.functional.binary_cross_entropy(theta.contiguous().view(theta.shape[0], -1),
                                                   priori_adj.contiguous().view(theta.shape[0], -1))
    loss_pred = masked_mae(preds=prediction, labels=real_value, null_val=null_val)
    return loss_pred + bce * gsl_coefficient
