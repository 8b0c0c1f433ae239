"""CPU: host-side mirror of the reference interface - constructor signatures, state-dict contract,
and the no-fallback rule (a CPU tensor must raise, never silently run in PyTorch)."""
import os

import pytest
import torch

from conftest import GOLDEN, ROOT, build_step_model
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


def test_scaler_is_loaded_and_applied_before_masking(tmp_path):
    """the runner must re-standardise prediction and label with the dataset's scaler pickle before the loss, so
    that a missing reading (raw 0 = -mean/std in z-score units) is masked (reference base_tsf_runner.py:36,238-250)."""
    import pickle
    import numpy as np
    from step.step_runner.scaler import load_scaler, rescale, scalar_stats
    from step.step_runner.metrics import masked_mae
    d = tmp_path / "datasets" / "X"
    d.mkdir(parents=True)
    cfg = {"TRAIN": {"DATA": {"DIR": str(d)}}, "DATASET_INPUT_LEN": 12, "DATASET_OUTPUT_LEN": 12}
    assert load_scaler(cfg) == {"mean": 0.0, "std": 1.0}                       # synthetic data: identity
    with open(d / "scaler_in12_out12.pkl", "wb") as f:
        pickle.dump({"func": "re_standard_transform", "args": {"mean": 54.4, "std": 19.5}}, f)
    sc = load_scaler(cfg)
    assert sc == {"mean": 54.4, "std": 19.5} and scalar_stats(sc) == (54.4, 19.5)
    label = torch.tensor([[-54.4 / 19.5, 1.0, 0.5]])                          # first entry: a missing reading
    pred = torch.tensor([[3.0, 1.0, 1.0]])
    got = float(masked_mae(rescale(pred, sc), rescale(label, sc), null_val=0.0))
    assert abs(got - 0.25 * 19.5) < 1e-4                                       # mean |err| over the two valid entries only
    wrong = float(masked_mae(pred, label, null_val=0.0))                       # z-score space: the missing entry is NOT masked
    assert abs(wrong - (3.0 + 54.4 / 19.5 + 0.5) / 3) < 1e-5
    arr = {"mean": np.array([1.0, 2.0], dtype=np.float32), "std": np.array([2.0, 3.0], dtype=np.float32)}
    assert torch.equal(rescale(torch.ones(4, 2), arr), torch.tensor([[3.0, 5.0]] * 4))
    assert load_scaler({"SCALER": {"mean": 1.0, "std": 2.0}}) == {"mean": 1.0, "std": 2.0}


def test_config_files_and_datasets_cover_all_six_datasets(tmp_path):
    import importlib
    from step.step_data import PretrainingDataset, ForecastingDataset
    for name in ("METR-LA", "PEMS-BAY", "PEMS03", "PEMS04", "PEMS07", "PEMS08"):
        s2 = importlib.import_module(f"step.STEP_{name}").CFG
        s1 = importlib.import_module(f"step.TSFormer_{name}").CFG
        assert s2.DATASET_CLS is ForecastingDataset and s1.DATASET_CLS is PretrainingDataset
        assert s2.MODEL.PARAM["dataset_name"] == name and s1.MODEL.PARAM["mode"] == "pre-train"
    assert importlib.import_module("step.TSFormer_PEMS03").CFG.DATASET_INPUT_LEN == 288 * 7      # shipped as-is (SURVEY Appx C)
    assert importlib.import_module("step.STEP_PEMS03").CFG.DATASET_ARGS["seq_len"] == 288 * 7 * 2
    ds = PretrainingDataset(synthetic=True, num_nodes=5, seq_len=48, length=6)
    future, history = ds[2]
    assert future.shape == (12, 5, 3) and history.shape == (48, 5, 3) and len(ds) == 6
    assert torch.equal(history, ds.data[2:50]) and torch.equal(future, ds.data[50:62])
    with pytest.raises(FileNotFoundError):
        PretrainingDataset(str(tmp_path / "nope.pkl"), str(tmp_path / "nope2.pkl"))


def test_run_py_never_touches_a_real_dataset_directory(tmp_path):
    """--synthetic writes into a scratch work directory and refuses to overwrite a data file it did not write."""
    import argparse
    import importlib.util
    spec = importlib.util.spec_from_file_location("run_mod", os.path.join(ROOT, "step", "run.py"))
    run = importlib.util.module_from_spec(spec); spec.loader.exec_module(run)
    cwd = os.getcwd()
    try:
        work = tmp_path / "w"
        args = argparse.Namespace(workdir=str(work), synthetic=True)
        run.enter_synthetic_workdir(args, "PEMS08", 170, 4032, need_ckpt=False)
        assert os.path.samefile(os.getcwd(), work) and os.path.isfile("datasets/PEMS08/data_in12_out12.pkl")
        assert not run.have_real_data("PEMS08", 12)                              # no index file -> not "real data"
        os.chdir(cwd)
        run.enter_synthetic_workdir(args, "PEMS08", 170, 4032, need_ckpt=False)   # re-running over its own files is fine
        os.chdir(cwd)
        real = tmp_path / "real" / "datasets" / "PEMS08"
        real.mkdir(parents=True)
        (real / "data_in12_out12.pkl").write_bytes(b"user data")
        with pytest.raises(SystemExit):
            run.enter_synthetic_workdir(argparse.Namespace(workdir=str(tmp_path / "real"), synthetic=True), "PEMS08", 170, 4032, False)
        assert (real / "data_in12_out12.pkl").read_bytes() == b"user data"
    finally:
        os.chdir(cwd)


def test_attention_dropout_threshold_drops_exactly_p_of_the_bf16_patterns():
    """The attention kernel decides two keep flags per 32-bit random word by comparing each 16-bit half, READ AS A bf16 NUMBER,
    with a threshold (one HSET2.BF16 per two probabilities).  Enumerate all 65536 patterns with the comparison semantics of
    the hardware (IEEE ordered >=: NaN is false, -0 == +0) and check that exactly floor(65536 p) patterns are dropped - the
    keep probability is exact although the patterns are compared as floats, not as integers."""
    import numpy as np
    from step_b200 import lib
    h = lib.load()
    pats = np.arange(65536, dtype=np.uint32)
    vals = (pats << 16).view(np.float32)                         # bf16 pattern -> its float value
    for p in (0.1, 0.05, 0.3, 0.25, 0.004, 0.45, 0.5):
        thr2 = int(h.step_tc_attn_drop_threshold(p))
        assert (thr2 & 0xFFFF) == (thr2 >> 16)                   # both halves carry the same threshold
        thr = np.array([(thr2 & 0xFFFF) << 16], dtype=np.uint32).view(np.float32)[0]
        with np.errstate(invalid="ignore"):
            kept = int(np.count_nonzero(vals >= thr))
        want_dropped = int(np.float32(p) * np.float32(65536.0))
        # +-0 compare equal: at p >= 0.498 the count can be off by the two zero patterns
        tol = 0 if want_dropped <= 254 + 32640 else 2
        assert abs((65536 - kept) - max(want_dropped, 254)) <= tol, (p, 65536 - kept, want_dropped)
    assert h.step_tc_attn_drop_threshold(0.0) == 0
