"""GPU tests of the training-loop pieces around the hot path: fused clip + Adam against
torch.nn.utils.clip_grad_norm_ + torch.optim.Adam, device-side metric accumulation against the metric functions,
easytorch-format checkpoint round trip, the device-resident window loader, and a few real optimisation steps."""
import math
import os

import pytest
import torch

from conftest import build_step_model
from oracle import step_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("max_norm,wd", [(3.0, 1e-5), (0.05, 0.0), (0.0, 1e-2)])
def test_fused_clip_adam_matches_torch(max_norm, wd):
    from step_b200.optim import FusedClipAdam
    g = torch.Generator().manual_seed(0)
    shapes = [(100, 40000), (32, 2, 1, 1), (7,), (256, 512), (3, 5, 17), (1,)]
    ref_p = [torch.nn.Parameter(torch.randn(*s, generator=g).to(DEV)) for s in shapes]
    my_p = [torch.nn.Parameter(p.detach().clone()) for p in ref_p]
    ref_opt = torch.optim.Adam(ref_p, lr=5e-3, weight_decay=wd, eps=1e-8)
    my_opt = FusedClipAdam(my_p, lr=5e-3, weight_decay=wd, eps=1e-8, max_norm=max_norm)
    for step in range(4):
        grads = [torch.randn(*s, generator=g).to(DEV) * (10.0 if step == 1 else 0.1) for s in shapes]
        skip = 2 if step == 2 else -1                      # a parameter without gradient is left untouched (as torch does)
        for i, (a, b, gr) in enumerate(zip(ref_p, my_p, grads)):
            a.grad = None if i == skip else gr.clone()
            b.grad = None if i == skip else gr.clone()
        if max_norm > 0:
            total = torch.nn.utils.clip_grad_norm_(ref_p, max_norm)
        else:
            total = torch.norm(torch.stack([p.grad.norm() for p in ref_p if p.grad is not None]))
        ref_opt.step()
        my_opt.step()
        assert abs(my_opt.grad_norm.item() - total.item()) < 1e-4 * max(1.0, total.item())
        for a, b in zip(ref_p, my_p):
            assert (a - b).abs().max().item() < 2e-6 * max(1.0, a.abs().max().item())
    # state dict interchange with torch.optim.Adam
    sd = my_opt.state_dict()
    ref2 = torch.optim.Adam([torch.nn.Parameter(p.detach().clone()) for p in my_p], lr=1.0)
    ref2.load_state_dict(sd)
    assert ref2.param_groups[0]["lr"] == 5e-3 and len(ref2.state) == len(shapes)
    again = FusedClipAdam([torch.nn.Parameter(p.detach().clone()) for p in my_p], lr=1.0)
    again.load_state_dict(ref_opt.state_dict())
    assert again.step_count == 4 and again.lr == 5e-3 and again._steps[0].tolist() == [4, 4, 3, 4, 4, 4]
    o = again.state_off[3]
    assert torch.allclose(again.exp_avg[o:o + 256 * 512].view(256, 512), ref_opt.state[ref_p[3]]["exp_avg"], atol=1e-7)


def test_metric_accumulator_matches_metric_functions():
    from step_b200.optim import MetricAccumulator
    from step.step_runner import metrics as M
    g = torch.Generator().manual_seed(1)
    acc = MetricAccumulator(DEV, null_val=0.0, mean=54.4, std=19.5)
    want = {"MAE": 0.0, "RMSE": 0.0, "MAPE": 0.0}
    for i in range(3):
        pred, real = torch.randn(4, 12, 50, 1, generator=g), torch.randn(4, 12, 50, 1, generator=g)
        real[torch.rand(real.shape, generator=g) < 0.2] = -54.4 / 19.5          # missing readings: raw value 0
        acc.update(pred.to(DEV), real.to(DEV))
        p, r = pred * 19.5 + 54.4, real * 19.5 + 54.4
        for k, f in M.METRICS.items():
            want[k] += float(f(p, r, null_val=0.0)) / 3
    got = acc.compute()
    assert got["batches"] == 3
    for k in want:
        assert abs(got[k] - want[k]) < 1e-4 * max(1.0, abs(want[k])), k
    acc.reset()
    assert acc.compute()["batches"] == 0


def test_checkpoint_round_trip_and_training_steps(tmp_path):
    """A few optimisation steps of the whole STEP model with the fused optimiser, an easytorch-format checkpoint, resume:
    the resumed run reproduces the continued one bit for bit (same seeds)."""
    from step.step_loss import step_loss
    from step.step_runner.checkpoint import load_checkpoint, save_checkpoint
    from step_b200.optim import FusedClipAdam
    ds = "PEMS08"
    history, long_history, future, uniform = O.synthetic_batch(ds, 2, 24, 0)
    history, long_history, future = history.to(DEV), long_history.to(DEV), future.to(DEV)

    def make():
        model, _, _ = build_step_model(tmp_path, ds, 0, real_ckpt=False)
        model = model.to(DEV).train()
        model.tsformer.dropout_p = 0.0
        model.backend.dropout = 0.0
        model.discrete_graph_learning.gumbel_uniform = uniform.to(DEV)
        opt = FusedClipAdam([p for p in model.parameters() if p.requires_grad], lr=2e-3, weight_decay=1e-5, max_norm=3.0)
        return model, opt

    def step(model, opt):
        y, theta, knn, coeff = model(history_data=history, long_history_data=long_history, future_data=None, batch_seen=0, epoch=1)
        loss = step_loss(y[..., [0]], future[..., [0]], theta, knn, coeff, null_val=0.0)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss.item()

    model, opt = make()
    losses = [step(model, opt) for _ in range(3)]
    assert all(math.isfinite(l) for l in losses) and losses[2] < losses[0]           # it learns on a fixed batch
    path = save_checkpoint(str(tmp_path / "ck" / "STEP_001.pt"), model, opt, 1, {"val_MAE": 3.5})
    ck = torch.load(path, map_location="cpu")
    assert set(ck) == {"epoch", "model_state_dict", "optim_state_dict", "best_metrics"} and ck["epoch"] == 1
    assert set(ck["optim_state_dict"]) == {"state", "param_groups"}
    cont = [step(model, opt) for _ in range(2)]
    model2, opt2 = make()
    info = load_checkpoint(path, model2, opt2)
    assert info == {"epoch": 1, "best_metrics": {"val_MAE": 3.5}} and opt2.step_count == 3
    resumed = [step(model2, opt2) for _ in range(2)]
    # same state, same inputs; split-K fp32 atomics make the weight gradients agree to ~1e-6 relative, not bitwise
    assert all(abs(a - b) < 1e-4 * max(1.0, abs(a)) for a, b in zip(resumed, cont))


def test_device_window_loader_on_gpu_feeds_the_runner():
    """DeviceWindowLoader with the series resident on the GPU: same items as the host dataset, and a training step consumes
    its batches without any host->device copy of the windows."""
    from step.step_data import DeviceWindowLoader, ForecastingDataset
    ds = ForecastingDataset(synthetic=True, num_nodes=11, seq_len=288, length=24, seed=2)
    loader = DeviceWindowLoader(ds, DEV, batch_size=8, shuffle=False)
    for bi, (future, history, long_history) in enumerate(loader):
        assert future.is_cuda and long_history.shape == (8, 288, 11, 3)
        ref = [ds[bi * 8 + i] for i in range(8)]
        assert torch.equal(future.cpu(), torch.stack([r[0] for r in ref]))
        assert torch.equal(history.cpu(), torch.stack([r[1] for r in ref]))
        assert torch.equal(long_history.cpu(), torch.stack([r[2] for r in ref]))
