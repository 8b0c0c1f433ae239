"""Runner glue for STEP: the `forward` / `train_iters` contract of the reference's STEPRunner
(step/step_runner/step_runner.py:7-75, basicts/runners/base_tsf_runner.py:225-255) without the easytorch training
loop it inherits from (out of scope here; easytorch is not vendored by the reference either).

`forward(data, epoch, iter_num, train)` takes the dataset tuple (future, history, long_history) on the HOST, moves it to the
running device, selects features, calls the model with the reference's keyword arguments and returns
(prediction, real_value, pred_adj, prior_adj, gsl_coefficient) - exactly what the reference hands to `cfg.TRAIN.LOSS`.
"""
import math

import torch

from .scaler import load_scaler, rescale


class STEPRunner:
    def __init__(self, cfg: dict, device=None):
        self.cfg = cfg
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.model = cfg["MODEL"]["ARCH"](**cfg["MODEL"]["PARAM"]).to(self.device)
        self.forward_features = cfg["MODEL"].get("FORWARD_FEATURES", None)
        self.target_features = cfg["MODEL"].get("TARGET_FEATURES", None)
        self.loss = cfg["TRAIN"]["LOSS"]
        self.null_val = cfg["TRAIN"].get("NULL_VAL", float("nan"))
        self.scaler = load_scaler(cfg)      # re_standard_transform arguments from the dataset's scaler pickle
        self.cl_param = cfg["TRAIN"].get("CL", None)
        self.output_seq_len = cfg.get("DATASET_OUTPUT_LEN", 12)
        self.iter_per_epoch = cfg.get("ITER_PER_EPOCH", 1)

    # ---- reference: step_runner.py:14-41 ----
    def select_input_features(self, data: torch.Tensor) -> torch.Tensor:
        if self.forward_features is not None and list(self.forward_features) != list(range(data.shape[-1])):
            data = data[:, :, :, self.forward_features]
        return data

    def select_target_features(self, data: torch.Tensor) -> torch.Tensor:
        return data[:, :, :, self.target_features]

    def to_running_device(self, t: torch.Tensor) -> torch.Tensor:
        return t.to(self.device, non_blocking=True)

    # ---- reference: step_runner.py:43-75 ----
    def forward(self, data: tuple, epoch: int = None, iter_num: int = None, train: bool = True, **kwargs) -> tuple:
        future_data, history_data, long_history_data = data
        history_data = self.to_running_device(history_data)
        long_history_data = self.to_running_device(long_history_data)
        future_data = self.to_running_device(future_data)
        history_data = self.select_input_features(history_data)
        long_history_data = self.select_input_features(long_history_data)
        prediction, pred_adj, prior_adj, gsl_coefficient = self.model(
            history_data=history_data, long_history_data=long_history_data, future_data=None, batch_seen=iter_num, epoch=epoch)
        batch_size, length, num_nodes, _ = future_data.shape
        assert list(prediction.shape)[:3] == [batch_size, length, num_nodes], \
            "error shape of the output, edit the forward function to reshape it to [B, L, N, C]"
        return self.select_target_features(prediction), self.select_target_features(future_data), pred_adj, prior_adj, gsl_coefficient

    # ---- reference: base_tsf_runner.py:225-255 (without the epoch meters) ----
    def rescale(self, x: torch.Tensor) -> torch.Tensor:
        return rescale(x, self.scaler)

    def curriculum_learning(self, epoch: int) -> int:
        if self.cl_param is None:
            return self.output_seq_len
        epoch -= 1
        warm, cl_epochs, step = self.cl_param["WARM_EPOCHS"], self.cl_param["CL_EPOCHS"], self.cl_param.get("STEP_SIZE", 1)
        if epoch < warm:
            return self.output_seq_len
        return min(math.ceil((epoch - warm + 1) / cl_epochs) * step, self.output_seq_len)

    def train_iters(self, epoch: int, iter_index: int, data: tuple) -> torch.Tensor:
        iter_num = (epoch - 1) * self.iter_per_epoch + iter_index
        ret = list(self.forward(data=data, epoch=epoch, iter_num=iter_num, train=True))
        pred, real = self.rescale(ret[0]), self.rescale(ret[1])
        if self.cl_param:
            cl = self.curriculum_learning(epoch)
            pred, real = pred[:, :cl], real[:, :cl]
        ret[0], ret[1] = pred, real
        return self.loss(*ret, null_val=self.null_val)

    # ---- reference: base_tsf_runner.py:275-318 (test): per-horizon and overall MAE / RMSE / MAPE on re-scaled values ----
    @torch.no_grad()
    def test(self, data_loader, horizons=None) -> dict:
        """data_loader yields the dataset tuples (future, history, long_history); the model is put in eval() for the
        pass (running BatchNorm statistics, no dropout, ``epoch=None`` -> gsl_coefficient 0) and restored afterwards."""
        from . import metrics
        was_training = self.model.training
        self.model.eval()
        try:
            pairs = []
            for data in data_loader:
                ret = self.forward(data=data, epoch=None, iter_num=None, train=False)
                pairs.append((self.rescale(ret[0]), self.rescale(ret[1])))
        finally:
            self.model.train(was_training)
        return metrics.evaluate(pairs, null_val=self.null_val, horizons=horizons)
