"""Golden values of the reference's masked metrics (basicts/metrics/{mae,rmse,mape}.py) on seeded inputs with missing
values - build container only:   python tests/golden/make_golden_metrics.py"""
import importlib.machinery
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def inputs(seed):
    g = torch.Generator().manual_seed(seed)
    real = torch.randn(5, 12, 9, 1, generator=g) * 20 + 50
    real[torch.rand(real.shape, generator=g) < 0.2] = 0.0            # missing sensor readings
    real[0, 0, 0, 0] = 3e-5                                           # inside the 5e-5 null tolerance
    pred = real + torch.randn(real.shape, generator=g) * 4
    return pred, real


def main():
    m = {n: importlib.machinery.SourceFileLoader("ref_" + n, os.path.join(REF, f"basicts/metrics/{n}.py")).load_module()
         for n in ("mae", "rmse", "mape")}
    fx = {}
    for seed in (0, 1):
        pred, real = inputs(seed)
        fx[seed] = {"MAE0": float(m["mae"].masked_mae(pred, real, 0.0)), "RMSE0": float(m["rmse"].masked_rmse(pred, real, 0.0)),
                    "MAPE": float(m["mape"].masked_mape(pred, real, 0.0)),
                    "MAEnan": float(m["mae"].masked_mae(pred, torch.where(real == 0, torch.full_like(real, float("nan")), real))),
                    "h3": [float(m["mae"].masked_mae(pred[:, 2], real[:, 2], 0.0)), float(m["rmse"].masked_rmse(pred[:, 2], real[:, 2], 0.0)),
                           float(m["mape"].masked_mape(pred[:, 2], real[:, 2], 0.0))]}
    torch.save(fx, os.path.join(HERE, "metrics_reference.pt"))
    print(fx)


if __name__ == "__main__":
    main()
