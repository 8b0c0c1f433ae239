"""Masked forecasting metrics with the semantics of basicts/metrics/{mae,rmse,mape}.py and the evaluation summary of
BaseTimeSeriesForecastingRunner.test (basicts/runners/base_tsf_runner.py:275-318): per-horizon and overall
MAE / RMSE / MAPE on the re-scaled predictions.

All three metrics share one weighting: labels within 5e-5 of ``null_val`` (or NaN labels when ``null_val`` is NaN) are
excluded, the remaining entries are weighted by 1 / (fraction kept) so that the mean over ALL entries equals the mean over
the kept ones, and NaNs produced by an empty selection count as zero.  Written as one masked-mean helper; works on any device.
"""
import math
from typing import Dict, Iterable, Optional, Sequence, Tuple

import torch


def _masked_mean(err: torch.Tensor, labels: torch.Tensor, null_val: float) -> torch.Tensor:
    keep = ~torch.isnan(labels) if math.isnan(null_val) else (labels - null_val).abs() > 5e-5
    w = keep.to(err.dtype)
    w = torch.nan_to_num(w / w.mean(), nan=0.0, posinf=0.0, neginf=0.0)
    return torch.nan_to_num(err * w, nan=0.0, posinf=float("inf"), neginf=float("-inf")).mean()


def masked_mae(preds: torch.Tensor, labels: torch.Tensor, null_val: float = float("nan")) -> torch.Tensor:
    return _masked_mean((preds - labels).abs(), labels, null_val)


def masked_mse(preds: torch.Tensor, labels: torch.Tensor, null_val: float = float("nan")) -> torch.Tensor:
    return _masked_mean((preds - labels) ** 2, labels, null_val)


def masked_rmse(preds: torch.Tensor, labels: torch.Tensor, null_val: float = float("nan")) -> torch.Tensor:
    return torch.sqrt(masked_mse(preds, labels, null_val))


def masked_mape(preds: torch.Tensor, labels: torch.Tensor, null_val: float = 0.0) -> torch.Tensor:
    """``null_val`` is fixed to 0 (as in the reference, mape.py:21): labels with |y| < 1e-4 are treated as missing."""
    labels = torch.where(labels.abs() < 1e-4, torch.zeros_like(labels), labels)
    return _masked_mean(((preds - labels).abs() / labels).abs(), labels, 0.0)


METRICS = {"MAE": masked_mae, "RMSE": masked_rmse, "MAPE": masked_mape}


@torch.no_grad()
def evaluate(batches: Iterable[Tuple[torch.Tensor, torch.Tensor]], scaler_mean: float = 0.0, scaler_std: float = 1.0,
             null_val: float = 0.0, horizons: Optional[Sequence[int]] = None) -> Dict[str, object]:
    """batches: iterable of (prediction [B,L,N,C], real_value [B,L,N,C]) in normalised units.  Returns
    {"horizon": {h: {metric: value}}, "overall": {metric: value}} on the re-scaled values (x * std + mean)."""
    preds, reals = zip(*[(p, r) for p, r in batches])
    pred = torch.cat(preds, 0) * scaler_std + scaler_mean
    real = torch.cat(reals, 0) * scaler_std + scaler_mean
    horizons = range(pred.shape[1]) if horizons is None else horizons
    out = {"horizon": {}, "overall": {}}
    for h in horizons:
        out["horizon"][h + 1] = {k: float(f(pred[:, h], real[:, h], null_val=null_val)) for k, f in METRICS.items()}
    out["overall"] = {k: float(f(pred, real, null_val=null_val)) for k, f in METRICS.items()}
    return out
