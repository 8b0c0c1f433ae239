"""easytorch-format checkpoints (SURVEY section 5 / Appx C): ``{"epoch", "model_state_dict", "optim_state_dict",
"best_metrics"}`` written by ``save_model(epoch)`` / ``save_best_model(epoch, "val_MAE", greater_best=False)``
(reference basicts/runners/base_runner.py:150, base_tsf_runner.py:320-329) and read by ``load_model(ckpt_path)``
(test/test_inference.py:21) and by ``STEP.load_pre_trained_model`` (step/step_arch/step.py:27-35).  File names follow
easytorch: ``<MODEL>_<epoch:03d>.pt`` and ``<MODEL>_best_<metric>.pt``."""
import os
from typing import Dict, Optional

import torch


def checkpoint_name(model_name: str, epoch: int) -> str:
    return "{0}_{1:03d}.pt".format(model_name, epoch)


def save_checkpoint(path: str, model: torch.nn.Module, optimizer=None, epoch: int = 0, best_metrics: Optional[Dict] = None) -> str:
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    ckpt = {"epoch": int(epoch), "model_state_dict": model.state_dict(),
            "optim_state_dict": None if optimizer is None else optimizer.state_dict(), "best_metrics": dict(best_metrics or {})}
    tmp = path + ".tmp"
    torch.save(ckpt, tmp)
    os.replace(tmp, path)               # a crash never leaves a truncated checkpoint behind
    return path


def load_checkpoint(path: str, model: torch.nn.Module, optimizer=None, strict: bool = True) -> Dict:
    """Loads on the CPU first (the shipped checkpoints were pickled from CUDA tensors) and lets
    ``load_state_dict`` copy onto the module's device."""
    ckpt = torch.load(path, map_location="cpu")
    model.load_state_dict(ckpt["model_state_dict"], strict=strict)
    if optimizer is not None and ckpt.get("optim_state_dict") is not None:
        optimizer.load_state_dict(ckpt["optim_state_dict"])
    return {"epoch": int(ckpt.get("epoch", 0)), "best_metrics": dict(ckpt.get("best_metrics") or {})}


class BestCheckpoint:
    """``save_best_model`` of the reference: keep the checkpoint of the epoch with the best validation metric."""

    def __init__(self, save_dir: str, model_name: str, metric: str = "val_MAE", greater_best: bool = False):
        self.save_dir, self.model_name, self.metric, self.greater_best = save_dir, model_name, metric, greater_best
        self.best_metrics: Dict[str, float] = {}

    def update(self, value: float, model, optimizer, epoch: int) -> bool:
        best = self.best_metrics.get(self.metric)
        better = best is None or (value > best if self.greater_best else value < best)
        if better:
            self.best_metrics[self.metric] = float(value)
            save_checkpoint(os.path.join(self.save_dir, "{0}_best_{1}.pt".format(self.model_name, self.metric)), model, optimizer,
                            epoch, self.best_metrics)
        return better
