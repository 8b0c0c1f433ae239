"""Runner glue for TSFormer pre-training (reference step/step_runner/tsformer_runner.py:43-71): the dataset tuple on the
host -> running device -> channel selection -> ``model(history_data=..., future_data=None, batch_seen, epoch)`` -> the
2-tuple (reconstruction of the masked patches, their ground truth) that the reference hands to ``cfg.TRAIN.LOSS``.

``train_iters`` returns the differentiable pre-training loss (the masked auto-encoder's backward runs on hand-written
kernels, ``TSFormer.pretrain_forward_autograd``); ``loss_iters`` evaluates it without a graph on the fused kernels."""
import torch

from .scaler import load_scaler, rescale


class TSFormerRunner:
    def __init__(self, cfg: dict, device=None):
        self.cfg = cfg
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.model = cfg["MODEL"]["ARCH"](**cfg["MODEL"]["PARAM"]).to(self.device)
        self.forward_features = cfg["MODEL"].get("FORWARD_FEATURES", None)
        self.loss = cfg["TRAIN"]["LOSS"]
        self.null_val = cfg["TRAIN"].get("NULL_VAL", float("nan"))
        self.scaler = load_scaler(cfg)
        self.iter_per_epoch = cfg.get("ITER_PER_EPOCH", 1)

    def select_input_features(self, data: torch.Tensor) -> torch.Tensor:
        return data if self.forward_features is None else data[:, :, :, self.forward_features]

    def forward(self, data: tuple, epoch: int = None, iter_num: int = None, train: bool = True, **kwargs) -> tuple:
        """data: (future, history) or the forecasting dataset's (future, history, long_history) - the LAST element is the
        long input window [B, P*12, N, C]."""
        history = self.select_input_features(data[-1].to(self.device, non_blocking=True))
        return self.model(history_data=history, future_data=None, batch_seen=iter_num, epoch=epoch)

    @torch.no_grad()
    def loss_iters(self, epoch: int, iter_index: int, data: tuple) -> torch.Tensor:
        """The pre-training objective on one batch (re-scaled like base_tsf_runner.py:238-250 does before the loss)."""
        rec, label = self.forward(data, epoch=epoch, iter_num=(epoch - 1) * self.iter_per_epoch + iter_index, train=False)
        return self.loss(rescale(rec, self.scaler), rescale(label, self.scaler), null_val=self.null_val)

    def train_iters(self, epoch: int, iter_index: int, data: tuple) -> torch.Tensor:
        """reference base_tsf_runner.py:225-255 for the 2-tuple model: forward, re-scale, ``cfg.TRAIN.LOSS``; the caller
        (easytorch's ``backward``) calls ``.backward()`` on the returned loss, clips and steps the optimiser."""
        rec, label = self.forward(data, epoch=epoch, iter_num=(epoch - 1) * self.iter_per_epoch + iter_index, train=True)
        return self.loss(rescale(rec, self.scaler), rescale(label, self.scaler), null_val=self.null_val)
