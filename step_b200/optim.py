"""Fused global-norm clip + Adam over all parameter tensors (two launches), and sync-free metric accumulation.

Drop-in for the optimiser part of the reference's training step (easytorch ``Runner.backward``: ``clip_grad_norm_(params,
max_norm)`` then ``torch.optim.Adam.step()``; configured at step/STEP_METR-LA.py:88-107).  Same update rule as
``torch.optim.Adam`` (L2 weight decay added to the gradient, bias correction, ``eps`` outside the square root), same
``state_dict()`` layout, so checkpoints interchange with the reference's ``optim_state_dict``."""
from __future__ import annotations

from typing import Dict, Iterable, List

import torch

from . import lib as _lib
from .lib import check


class FusedClipAdam:
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, max_norm: float = 0.0):
        self.params: List[torch.nn.Parameter] = [p for p in params]
        if not self.params:
            raise ValueError("FusedClipAdam: no parameters")
        for p in self.params:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise _lib.StepB200Error("FusedClipAdam: parameters must be contiguous float32 CUDA tensors (no CPU path)")
        self.lr, self.betas, self.eps, self.weight_decay, self.max_norm = float(lr), tuple(betas), float(eps), float(weight_decay), float(max_norm)
        self.step_count = 0
        dev = self.params[0].device
        self.device = dev
        L = _lib.load()
        chunk = int(L.step_opt_chunk_elems())
        numel = [p.numel() for p in self.params]
        offs, tot = [], 0
        for n in numel:
            offs.append(tot)
            tot += (n + 3) // 4 * 4
        self.state_off = offs
        self.exp_avg = torch.zeros(tot, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(tot, device=dev, dtype=torch.float32)
        ct, co = [], []
        for i, n in enumerate(numel):
            for o in range(0, n, chunk):
                ct.append(i)
                co.append(o)
        self.n_chunks = len(ct)
        i64 = lambda x: torch.tensor(x, dtype=torch.int64, device=dev)
        self._p_ptr = i64([p.data_ptr() for p in self.params])
        self._numel, self._soff = i64(numel), i64(offs)
        self._chunk_tensor = torch.tensor(ct, dtype=torch.int32, device=dev)
        self._chunk_off = i64(co)
        self._g_host = torch.zeros(len(self.params), dtype=torch.int64).pin_memory()
        self._g_dev = torch.zeros(len(self.params), dtype=torch.int64, device=dev)
        self._sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
        self._steps = [torch.zeros(len(self.params), dtype=torch.int32, device=dev) for _ in range(2)]   # ping-pong: in / out
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=dev)      # norm before clipping, of the last step (device)

    def zero_grad(self) -> None:
        for p in self.params:
            p.grad = None                      # autograd then ASSIGNS the new gradient: no accumulate kernel, no memset

    @torch.no_grad()
    def step(self) -> None:
        self.step_count += 1
        keep = []
        for i, p in enumerate(self.params):
            g = p.grad
            if g is not None and not g.is_contiguous():
                g = g.contiguous()
                p.grad = g
            if g is not None and (g.dtype != torch.float32 or g.device != p.device):
                raise _lib.StepB200Error("FusedClipAdam: gradients must be float32 on the parameter's device")
            keep.append(g)
            self._g_host[i] = 0 if g is None else g.data_ptr()
        self._g_dev.copy_(self._g_host, non_blocking=True)
        L = _lib.load()
        check(L.step_set_device(self.device.index), "step_set_device")
        st = torch.cuda.current_stream(self.device).cuda_stream
        check(L.step_clip_adam_step(self._p_ptr.data_ptr(), self._g_dev.data_ptr(), self._numel.data_ptr(), self._soff.data_ptr(),
                                    self._chunk_tensor.data_ptr(), self._chunk_off.data_ptr(), self.n_chunks,
                                    self._steps[0].data_ptr(), self._steps[1].data_ptr(),
                                    self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self._sumsq.data_ptr(), self.max_norm,
                                    self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                    self.grad_norm.data_ptr(), st), "step_clip_adam_step")
        self._steps.reverse()                      # the kernel wrote the advanced per-tensor step counts into the second buffer

    # ---- torch.optim.Adam-compatible state dict (easytorch checkpoints store optimizer.state_dict()) ----
    def state_dict(self) -> Dict:
        state = {}
        steps = self._steps[0].tolist()
        for i, p in enumerate(self.params):
            if steps[i] == 0:
                continue                       # torch creates a parameter's state at its first step with a gradient
            o, n = self.state_off[i], p.numel()
            state[i] = {"step": torch.tensor(float(steps[i])), "exp_avg": self.exp_avg[o:o + n].view_as(p).clone(),
                        "exp_avg_sq": self.exp_avg_sq[o:o + n].view_as(p).clone()}
        group = {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd: Dict) -> None:
        group = sd["param_groups"][0]
        if len(group["params"]) != len(self.params):
            raise ValueError(f"FusedClipAdam: checkpoint has {len(group['params'])} parameters, the model {len(self.params)}")
        self.lr, self.betas, self.eps, self.weight_decay = float(group["lr"]), tuple(group["betas"]), float(group["eps"]), float(group["weight_decay"])
        steps = [0] * len(self.params)
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        for i, pid in enumerate(group["params"]):
            s = sd["state"].get(pid)
            if s is None:
                continue
            steps[i] = int(float(s["step"]))
            o, n = self.state_off[i], self.params[i].numel()
            self.exp_avg[o:o + n].copy_(s["exp_avg"].reshape(-1).to(self.device))
            self.exp_avg_sq[o:o + n].copy_(s["exp_avg_sq"].reshape(-1).to(self.device))
        self.step_count = max(steps) if steps else 0
        self._steps[0].copy_(torch.tensor(steps, dtype=torch.int32))


class MetricAccumulator:
    """Epoch meters for masked MAE / RMSE / MAPE kept on the device: ``update`` enqueues two kernels and never syncs;
    ``compute`` (once per epoch / log line) reads the averages of the per-batch values, which is what the reference's
    ``update_epoch_meter(..., metric.item())`` averages with three host syncs per step (base_tsf_runner.py:252-254)."""

    def __init__(self, device, null_val: float = float("nan"), mean: float = 0.0, std: float = 1.0):
        import math
        self.device = torch.device(device)
        self.nan_mask = 1 if (isinstance(null_val, float) and math.isnan(null_val)) else 0
        self.null_val = 0.0 if self.nan_mask else float(null_val)
        self.mean, self.std = float(mean), float(std)
        self._sums = torch.zeros(5, dtype=torch.float64, device=self.device)
        self._acc = torch.zeros(4, dtype=torch.float64, device=self.device)

    @torch.no_grad()
    def update(self, pred: torch.Tensor, real: torch.Tensor) -> None:
        pred, real = pred.detach().contiguous(), real.detach().contiguous()
        if not pred.is_cuda or pred.dtype != torch.float32 or pred.shape != real.shape:
            raise _lib.StepB200Error("MetricAccumulator: float32 CUDA tensors of equal shape expected (no CPU path)")
        L = _lib.load()
        check(L.step_set_device(self.device.index), "step_set_device")
        st = torch.cuda.current_stream(self.device).cuda_stream
        check(L.step_metrics_accumulate(pred.data_ptr(), real.data_ptr(), pred.numel(), self.mean, self.std, self.null_val,
                                        self.nan_mask, self._sums.data_ptr(), self._acc.data_ptr(), st), "step_metrics_accumulate")

    def compute(self) -> Dict[str, float]:
        a = self._acc.cpu()
        n = max(float(a[3]), 1.0)
        return {"MAE": float(a[0]) / n, "RMSE": float(a[1]) / n, "MAPE": float(a[2]) / n, "batches": int(a[3])}

    def reset(self) -> None:
        self._acc.zero_()
