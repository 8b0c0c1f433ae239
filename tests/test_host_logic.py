"""CPU: host-side mirror of the reference interface - constructor signatures, state-dict contract,
and the no-fallback rule (a CPU tensor must raise, never silently run in PyTorch)."""
import os

import pytest
import torch

from conftest import GOLDEN, build_step_model
from oracle import step_oracle as O


def expected_keys(dataset):
    keys = set(O.synthetic_trainable_params(dataset).keys()) | set(O.bn_buffers(dataset).keys())
    keys |= {"tsformer." + k for k in O.synthetic_tsformer_params().keys()}
    return keys


def test_state_dict_contract_and_ckpt_load(tmp_path):
    model, full, _ = build_step_model(tmp_path, "PEMS08")
    assert set(model.state_dict().keys()) == expected_keys("PEMS08")
    for k, v in model.state_dict().items():
        assert tuple(v.shape) == tuple(full[k].shape), k
    # frozen TSFormer, trainable rest (reference step.py:34-35)
    assert not any(p.requires_grad for p in model.tsformer.parameters())
    assert all(p.requires_grad for p in model.backend.parameters())
    # the real METR-LA checkpoint's 72 keys load strictly
    real = torch.load(os.path.join(GOLDEN, "tsformer_METR-LA_state.pt"))
    assert len(real) == 72
    model.tsformer.load_state_dict(real, strict=True)


def test_no_cpu_fallback(tmp_path):
    from step_b200.lib import StepB200Error
    model, _, _ = build_step_model(tmp_path, "PEMS08")
    history, long_history, _, _ = O.synthetic_batch("PEMS08", 1, 2, 0)
    with pytest.raises(StepB200Error):
        model(history_data=history, long_history_data=long_history, future_data=None, batch_seen=0, epoch=1)


def test_step_loss_matches_oracle():
    from step.step_loss import step_loss
    g = torch.Generator().manual_seed(0)
    pred, real = torch.randn(3, 12, 11, 1, generator=g), torch.randn(3, 12, 11, 1, generator=g)
    real[0, :, 3] = 0.0
    theta = torch.rand(11, 11, generator=g).unsqueeze(0).expand(3, 11, 11)
    prior = (torch.rand(3, 11, 11, generator=g) > 0.9).float()
    a = step_loss(pred, real, theta, prior, 0.5, null_val=0.0)
    b = O.step_loss(pred, real, theta, prior, 0.5, null_val=0.0)
    assert abs(a.item() - b.item()) < 1e-6


def test_mask_generator_reproduces_the_reference_draw():
    """MaskGenerator (reference tsformer/mask.py:15-24) consumes Python's ``random`` exactly like the reference: with
    the seed the golden fixture was generated under, it returns the mask the reference drew."""
    import random
    from step.step_arch.tsformer.tsformer import MaskGenerator
    fx = torch.load(os.path.join(GOLDEN, "tsformer_pretrain_METR-LA.pt"), weights_only=False)
    mg = MaskGenerator(float(fx["P"]), 0.75)          # num_token arrives as a float in the shipped configs
    random.seed(fx["random_seed"])
    unmasked, masked = mg()
    assert unmasked == fx["unmasked"] and masked == fx["masked"]
    mg.fixed = ([1, 2], [0, 3])
    assert mg() == ([1, 2], [0, 3])


def test_device_window_loader_matches_dataset_items():
    """DeviceWindowLoader (device-resident series, windows gathered from the index) serves exactly the items of
    ForecastingDataset.__getitem__ stacked over the batch, including the all-zero long history of early windows."""
    from step.step_data import DeviceWindowLoader, ForecastingDataset
    ds = ForecastingDataset(synthetic=True, num_nodes=7, seq_len=48, length=40, seed=3)
    ds.index = [(i, i + 12, i + 24) for i in range(20, 60)]          # the first windows start before seq_len
    loader = DeviceWindowLoader(ds, "cpu", batch_size=16, shuffle=True, seed=5)
    assert len(loader) == 3
    seen = 0
    g = torch.Generator().manual_seed(5)
    order = torch.randperm(40, generator=g)
    for bi, (future, history, long_history) in enumerate(loader):
        ids = order[bi * 16:(bi + 1) * 16]
        ref = [ds[int(i)] for i in ids]
        assert torch.equal(future, torch.stack([r[0] for r in ref]))
        assert torch.equal(history, torch.stack([r[1] for r in ref]))
        assert torch.equal(long_history, torch.stack([r[2] for r in ref]))
        seen += len(ids)
    assert seen == 40
    assert any(float(ds[i][2].abs().sum()) == 0.0 for i in range(5))          # the zero-window branch was exercised
    only0 = DeviceWindowLoader(ds, "cpu", batch_size=8, long_channels=[0])
    f, h, lh = next(iter(only0))
    assert lh.shape == (8, 48, 7, 1) and torch.equal(lh[..., 0], torch.stack([ds[i][2][..., 0] for i in range(8)]))


def test_metrics_match_reference_golden():
    """step_runner.metrics (MAE / RMSE / MAPE with missing values, per-horizon summary) vs values computed by the
    reference's own metric functions (tests/golden/make_golden_metrics.py)."""
    import importlib.util
    from step.step_runner import metrics as M
    spec = importlib.util.spec_from_file_location("mgm", os.path.join(GOLDEN, "make_golden_metrics.py"))
    mgm = importlib.util.module_from_spec(spec); spec.loader.exec_module(mgm)
    fx = torch.load(os.path.join(GOLDEN, "metrics_reference.pt"), weights_only=False)
    for seed, want in fx.items():
        pred, real = mgm.inputs(seed)
        assert abs(float(M.masked_mae(pred, real, 0.0)) - want["MAE0"]) < 1e-5 * max(1.0, abs(want["MAE0"]))
        assert abs(float(M.masked_rmse(pred, real, 0.0)) - want["RMSE0"]) < 1e-5 * max(1.0, abs(want["RMSE0"]))
        assert abs(float(M.masked_mape(pred, real, 0.0)) - want["MAPE"]) < 1e-5 * max(1.0, abs(want["MAPE"]))
        nan_real = torch.where(real == 0, torch.full_like(real, float("nan")), real)
        assert abs(float(M.masked_mae(pred, nan_real)) - want["MAEnan"]) < 1e-5 * max(1.0, abs(want["MAEnan"]))
        rep = M.evaluate([(pred[:2], real[:2]), (pred[2:], real[2:])], null_val=0.0)
        got = rep["horizon"][3]
        for v, w in zip((got["MAE"], got["RMSE"], got["MAPE"]), want["h3"]):
            assert abs(v - w) < 1e-5 * max(1.0, abs(w))
        assert abs(rep["overall"]["MAE"] - want["MAE0"]) < 1e-5 * max(1.0, abs(want["MAE0"]))
    # an all-missing slice contributes zero instead of NaN
    z = torch.zeros(2, 3)
    assert float(M.masked_mae(torch.ones(2, 3), z, 0.0)) == 0.0


def test_tsformer_pretrain_config_and_runner_contract():
    """Stage-1 config layout (reference step/TSFormer_<NAME>.py) and the TSFormerRunner glue: constructor arguments the
    reference passes, 72 checkpoint keys, masked-MAE objective, no CPU path."""
    import importlib
    from step.step_runner import TSFormerRunner
    from step.step_runner.metrics import masked_mae
    for name, batch, lr, tokens in (("METR-LA", 8, 0.0005, 168.0), ("PEMS04", 6, 0.001, 336.0)):
        CFG = importlib.import_module(f"step.TSFormer_{name}").CFG
        assert CFG.RUNNER is TSFormerRunner and CFG.TRAIN.LOSS is masked_mae and CFG.TRAIN.NULL_VAL == 0.0
        assert CFG.TRAIN.DATA.BATCH_SIZE == batch and CFG.TRAIN.OPTIM.PARAM["lr"] == lr
        assert CFG.MODEL.PARAM["mode"] == "pre-train" and CFG.MODEL.PARAM["num_token"] == tokens
        assert CFG.MODEL.FORWARD_FEATURES == [0] and CFG.DATASET_INPUT_LEN == int(tokens) * 12
    runner = TSFormerRunner(CFG, device="cpu")
    assert len(runner.model.state_dict()) == 72 and runner.model.mode == "pre-train"
    from step_b200.lib import StepB200Error
    with pytest.raises(StepB200Error):                     # stage-1 training exists on the GPU only: CPU tensors fail loudly
        runner.train_iters(1, 0, (torch.zeros(1, 12, 3, 3), torch.zeros(1, 4032, 3, 3)))
    with pytest.raises(StepB200Error):                     # CPU tensors never run silently
        runner.loss_iters(1, 0, (torch.zeros(1, 12, 3, 3), torch.zeros(1, 4032, 3, 3)))


def test_step_forward_orchestration_with_stub_submodules(tmp_path):
    """STEP.forward wiring (reference step.py:37-72) checked on CPU with stub sub-modules: last-patch hidden state to the
    backend, [B,N,12] -> [B,12,N,1], theta as a stride-0 [B,N,N] view, gsl coefficient 1/(epoch//6+1) or 0."""
    model, _, _ = build_step_model(tmp_path, "PEMS08")
    B, N, P = 2, 170, 3
    hidden = torch.randn(B, N, P, 96)
    seen = {}

    class FakeDGL(torch.nn.Module):
        theta = torch.rand(N, N)

        def forward(self, long_history, tsformer):
            seen["long"] = long_history.shape
            return torch.zeros(B, N * N, 2), hidden, torch.ones(B, N, N), torch.zeros(B, N, N)

    class FakeBackend(torch.nn.Module):
        def forward(self, x, hidden_states, sampled_adj):
            seen["hidden_last"] = hidden_states
            return torch.arange(B * N * 12, dtype=torch.float32).view(B, N, 12)

    model.discrete_graph_learning, model.backend = FakeDGL(), FakeBackend()
    history, long_history = torch.zeros(B, 12, N, 3), torch.zeros(B, P * 12, N, 3)
    y, theta, knn, coeff = model(history_data=history, long_history_data=long_history, future_data=None, batch_seen=0, epoch=13)
    assert torch.equal(seen["hidden_last"], hidden[:, :, -1, :]) and tuple(seen["long"]) == (B, P * 12, N, 3)
    assert y.shape == (B, 12, N, 1) and float(y[1, 3, 5, 0]) == float((1 * N + 5) * 12 + 3)
    assert theta.shape == (B, N, N) and theta.stride(0) == 0 and torch.equal(theta[1], FakeDGL.theta)
    assert coeff == 1 / 3 and knn.shape == (B, N, N)
    assert model(history_data=history, long_history_data=long_history, future_data=None, batch_seen=0, epoch=None)[3] == 0
