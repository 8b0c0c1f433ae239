"""Re-normalisation of predictions and labels before the loss / metrics.

Reference: basicts/runners/base_tsf_runner.py:36 loads ``{TRAIN.DATA.DIR}/scaler_in{IN}_out{OUT}.pkl`` =
``{"func": "re_standard_transform", "args": {"mean": ..., "std": ...}}`` and applies ``x * std + mean``
(basicts/data/transform.py:48-65) to both tensors before ``cfg.TRAIN.LOSS`` (base_tsf_runner.py:238-250).  Without it the
null-value mask would be taken in z-score space, where a missing reading (raw 0) is ``-mean/std`` and is NOT masked."""
import os
import pickle

import numpy as np
import torch


def load_scaler(cfg) -> dict:
    """{"mean": float | Tensor, "std": float | Tensor}.  Order: an explicit ``cfg.SCALER``; the scaler pickle of the dataset
    directory; identity (synthetic data has no scaler file)."""
    if cfg.get("SCALER") is not None:
        return dict(cfg["SCALER"])
    try:
        path = "{0}/scaler_in{1}_out{2}.pkl".format(cfg["TRAIN"]["DATA"]["DIR"], cfg["DATASET_INPUT_LEN"], cfg["DATASET_OUTPUT_LEN"])
    except (KeyError, TypeError):
        path = None
    if path and os.path.isfile(path):
        with open(path, "rb") as f:
            try:
                sc = pickle.load(f)
            except UnicodeDecodeError:
                f.seek(0)
                sc = pickle.load(f, encoding="latin1")
        func = sc.get("func", "re_standard_transform")
        if func != "re_standard_transform":
            raise NotImplementedError(f"scaler function {func!r}: every STEP dataset uses re_standard_transform")
        return {"mean": sc["args"]["mean"], "std": sc["args"]["std"]}
    return {"mean": 0.0, "std": 1.0}


def rescale(x: torch.Tensor, scaler: dict) -> torch.Tensor:
    """re_standard_transform (transform.py:48-65): ndarray statistics are broadcast over the batch axis."""
    mean, std = scaler["mean"], scaler["std"]
    if isinstance(mean, np.ndarray):
        mean = torch.from_numpy(mean).type_as(x).to(x.device).unsqueeze(0)
        std = torch.from_numpy(std).type_as(x).to(x.device).unsqueeze(0)
    return x * std + mean


def scalar_stats(scaler: dict):
    """(mean, std) as Python floats when the statistics are scalars (the fused loss kernel takes them as arguments)."""
    m, s = scaler["mean"], scaler["std"]
    if isinstance(m, (int, float)) and isinstance(s, (int, float)):
        return float(m), float(s)
    if isinstance(m, np.ndarray) and m.size == 1 and isinstance(s, np.ndarray) and s.size == 1:
        return float(m.reshape(-1)[0]), float(s.reshape(-1)[0])
    return None
