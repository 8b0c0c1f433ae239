"""`python step/run.py --cfg step/STEP_METR-LA.py --gpus 0` - same CLI as the reference's step/run.py:13-33.

The reference hands the config to easytorch's training loop (not vendored, out of scope).  This entry point builds the
same runner/model from the same config layout and drives `train_iters` + Adam + grad clipping for `--steps` iterations;
without dataset files (none are shipped, no network) it writes a synthetic `datasets/<NAME>/data_in12_out12.pkl` and
random-initialises the TSFormer checkpoint so that the whole path can be exercised end to end on a GPU."""
import importlib
import os
import pickle
import sys
import time
from argparse import ArgumentParser

sys.path.append(os.path.abspath(__file__ + "/../.."))
import torch  # noqa: E402


def parse_args():
    parser = ArgumentParser(description="Run STEP on the B200-native kernels")
    parser.add_argument("-c", "--cfg", default="step/STEP_METR-LA.py", help="training config")
    parser.add_argument("--gpus", default="0", help="visible gpus")
    parser.add_argument("--steps", type=int, default=20)
    parser.add_argument("--synthetic", action="store_true", help="force synthetic data / random TSFormer weights")
    return parser.parse_args()


def pretrain_forward(args, mod):
    """`--cfg step/TSFormer_<NAME>.py`: stage-1 config.  The masked encoder/decoder run forward-only on the B200 kernels
    (DESIGN.md section 8), so this evaluates the pre-training objective on synthetic windows instead of training."""
    from step.step_data import ForecastingDataset
    CFG = importlib.import_module(mod).CFG
    name = CFG.DATASET_NAME
    from step.configs import _NODES
    torch.manual_seed(CFG.ENV.SEED)
    runner = CFG.RUNNER(CFG)
    runner.model.train()                                   # dropout live, as in the reference's training iterations
    ds = ForecastingDataset(mode="train", seq_len=CFG.DATASET_INPUT_LEN, synthetic=True, num_nodes=_NODES[name],
                            length=CFG.TRAIN.DATA.BATCH_SIZE * 4)
    loader = torch.utils.data.DataLoader(ds, batch_size=CFG.TRAIN.DATA.BATCH_SIZE, shuffle=True, drop_last=True, pin_memory=True)
    it, t0 = 0, time.perf_counter()
    while it < args.steps:
        for data in loader:
            loss = runner.loss_iters(1, it, data)
            it += 1
            if it % 5 == 0 or it == args.steps:
                torch.cuda.synchronize()
                print(f"iter {it:4d}  reconstruction MAE {loss.item():.5f}  "
                      f"{it * CFG.TRAIN.DATA.BATCH_SIZE / (time.perf_counter() - t0):.1f} samples/s (forward only)")
            if it >= args.steps:
                break


def main():
    args = parse_args()
    os.environ.setdefault("CUDA_VISIBLE_DEVICES", args.gpus)
    mod = "step." + os.path.splitext(os.path.basename(args.cfg))[0]
    from step.step_data import ForecastingDataset
    if os.path.basename(args.cfg).startswith("TSFormer_"):
        return pretrain_forward(args, mod)
    name = os.path.basename(args.cfg)[len("STEP_"):-3]
    from step.configs import _NODES, _SEQ
    n, seq = _NODES[name], _SEQ[name]
    data_dir = os.path.join("datasets", name)
    have_data = os.path.isfile(os.path.join(data_dir, "data_in12_out12.pkl")) and not args.synthetic
    if not have_data:
        os.makedirs(data_dir, exist_ok=True)
        g = torch.Generator().manual_seed(0)
        with open(os.path.join(data_dir, "data_in12_out12.pkl"), "wb") as f:
            pickle.dump({"processed_data": torch.randn(40000, n, 3, generator=g).numpy()}, f)
    ckpt = os.path.join("tsformer_ckpt", f"TSFormer_{name}.pt")
    if not os.path.isfile(ckpt):
        from step.step_arch import TSFormer
        os.makedirs("tsformer_ckpt", exist_ok=True)
        ts = TSFormer(patch_size=12, in_channel=1, embed_dim=96, num_heads=4, mlp_ratio=4, dropout=0.1, num_token=seq / 12,
                      mask_ratio=0.75, encoder_depth=4, decoder_depth=1, mode="forecasting")
        torch.save({"model_state_dict": ts.state_dict()}, ckpt)
    CFG = importlib.import_module(mod).CFG
    torch.manual_seed(CFG.ENV.SEED)
    runner = CFG.RUNNER(CFG)
    runner.model.train()
    opt = torch.optim.Adam([p for p in runner.model.parameters() if p.requires_grad], **CFG.TRAIN.OPTIM.PARAM)
    if have_data:
        ds = ForecastingDataset(os.path.join(data_dir, "data_in12_out12.pkl"), os.path.join(data_dir, "index_in12_out12.pkl"),
                                "train", seq)
    else:
        ds = ForecastingDataset(mode="train", seq_len=seq, synthetic=True, num_nodes=n, length=CFG.TRAIN.DATA.BATCH_SIZE * 4)
    loader = torch.utils.data.DataLoader(ds, batch_size=CFG.TRAIN.DATA.BATCH_SIZE, shuffle=True, drop_last=True, pin_memory=True)
    it, t0 = 0, time.perf_counter()
    while it < args.steps:
        for data in loader:
            loss = runner.train_iters(1, it, data)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(runner.model.parameters(), **CFG.TRAIN.CLIP_GRAD_PARAM)
            opt.step()
            it += 1
            if it % 5 == 0 or it == args.steps:
                torch.cuda.synchronize()
                print(f"iter {it:4d}  loss {loss.item():.5f}  {it * CFG.TRAIN.DATA.BATCH_SIZE / (time.perf_counter() - t0):.1f} samples/s")
            if it >= args.steps:
                break


if __name__ == "__main__":
    main()
