"""`python step/run.py --cfg step/STEP_METR-LA.py --gpus 0` - same CLI as the reference's step/run.py:13-33.

The reference hands the config to easytorch's training loop (not vendored, out of scope).  This entry point builds the
same runner / model from the same config layout and drives `train_iters` + the fused clip+Adam step for `--steps`
iterations, for both stages: `step/TSFormer_<NAME>.py` (masked auto-encoder pre-training) and `step/STEP_<NAME>.py`.

Data: with `datasets/<NAME>/{data,index,scaler}_in*_out12.pkl` present (the reference's pre-processing output) they are
used as they are and NEVER written to.  Otherwise - or with `--synthetic` - a seeded synthetic series and a randomly
initialised TSFormer checkpoint are written into a fresh temporary directory (`--workdir`, default a mkdtemp), and the run
happens there: the dataset directory of a real checkout is not touched."""
import importlib
import os
import pickle
import sys
import tempfile
import time
from argparse import ArgumentParser

sys.path.append(os.path.abspath(__file__ + "/../.."))
import torch  # noqa: E402


def parse_args():
    parser = ArgumentParser(description="Run STEP on the B200-native kernels")
    parser.add_argument("-c", "--cfg", default="step/STEP_METR-LA.py", help="training config")
    parser.add_argument("--gpus", default="0", help="visible gpus")
    parser.add_argument("--steps", type=int, default=20)
    parser.add_argument("--synthetic", action="store_true", help="force synthetic data / random TSFormer weights (in --workdir)")
    parser.add_argument("--workdir", default=None, help="where synthetic data and checkpoints are written (default: a temp dir)")
    parser.add_argument("--device-loader", action="store_true",
                        help="stage 2: keep the series resident on the GPU and gather every batch there (DeviceWindowLoader)")
    parser.add_argument("--save", default=None, help="write an easytorch-format checkpoint here after the last step")
    parser.add_argument("--resume", default=None, help="resume model + optimiser state from this checkpoint")
    return parser.parse_args()


def dataset_files(name, in_len):
    d = os.path.join("datasets", name)
    return (os.path.join(d, "data_in{0}_out12.pkl".format(in_len)), os.path.join(d, "index_in{0}_out12.pkl".format(in_len)))


def have_real_data(name, in_len):
    return all(os.path.isfile(p) for p in dataset_files(name, in_len))


def enter_synthetic_workdir(args, name, nodes, seq, need_ckpt):
    """chdir into a scratch directory holding `datasets/<NAME>/data_in12_out12.pkl` (the discrete-graph-learning module reads
    it relative to the CWD, reference discrete_graph_learning.py:57) and, for stage 2, a random TSFormer checkpoint."""
    work = args.workdir or tempfile.mkdtemp(prefix="step_run_")
    work = os.path.abspath(work)
    ddir = os.path.join(work, "datasets", name)
    os.makedirs(ddir, exist_ok=True)
    pkl = os.path.join(ddir, "data_in12_out12.pkl")
    marker = os.path.join(ddir, "SYNTHETIC")
    if os.path.exists(pkl) and not os.path.exists(marker):
        raise SystemExit(f"run.py: {pkl} exists and was not written by --synthetic; refusing to overwrite it")
    g = torch.Generator().manual_seed(0)
    with open(pkl, "wb") as f:
        pickle.dump({"processed_data": torch.randn(40000, nodes, 3, generator=g).numpy()}, f)
    open(marker, "w").write("written by step/run.py --synthetic\n")
    if need_ckpt:
        ck = os.path.join(work, "tsformer_ckpt", f"TSFormer_{name}.pt")
        if not os.path.isfile(ck):
            from step.step_arch import TSFormer
            os.makedirs(os.path.dirname(ck), exist_ok=True)
            ts = TSFormer(patch_size=12, in_channel=1, embed_dim=96, num_heads=4, mlp_ratio=4, dropout=0.1, num_token=seq / 12,
                          mask_ratio=0.75, encoder_depth=4, decoder_depth=1, mode="forecasting")
            torch.save({"model_state_dict": ts.state_dict()}, ck)
    os.chdir(work)
    print(f"run.py: synthetic data and checkpoints under {work}")


def main():
    args = parse_args()
    os.environ.setdefault("CUDA_VISIBLE_DEVICES", args.gpus)
    base = os.path.splitext(os.path.basename(args.cfg))[0]
    stage1 = base.startswith("TSFormer_")
    name = base[len("TSFormer_"):] if stage1 else base[len("STEP_"):]
    from step.configs import _NODES, _SEQ, _TS_SEQ
    from step.step_data import ForecastingDataset, PretrainingDataset
    from step.step_runner.checkpoint import load_checkpoint, save_checkpoint
    from step_b200.optim import FusedClipAdam
    nodes = _NODES[name]
    seq = _TS_SEQ[name] if stage1 else _SEQ[name]
    in_len = seq if stage1 else 12
    real = have_real_data(name, in_len) and not args.synthetic
    if real and not stage1 and not os.path.isfile(os.path.join("tsformer_ckpt", f"TSFormer_{name}.pt")):
        raise SystemExit(f"run.py: tsformer_ckpt/TSFormer_{name}.pt is missing (pre-train it with step/TSFormer_{name}.py)")
    if not real:
        enter_synthetic_workdir(args, name, nodes, seq, need_ckpt=not stage1)
    CFG = importlib.import_module("step." + base).CFG
    torch.manual_seed(CFG.ENV.SEED)
    runner = CFG.RUNNER(CFG)
    runner.model.train()
    batch = CFG.TRAIN.DATA.BATCH_SIZE
    if real:
        data_file, index_file = dataset_files(name, in_len)
        ds = PretrainingDataset(data_file, index_file, "train") if stage1 else ForecastingDataset(data_file, index_file, "train", seq)
    elif stage1:
        ds = PretrainingDataset(mode="train", synthetic=True, num_nodes=nodes, seq_len=seq, length=batch * 4)
    else:
        ds = ForecastingDataset(mode="train", seq_len=seq, synthetic=True, num_nodes=nodes, length=batch * 4)
    if args.device_loader and not stage1:
        from step.step_data import DeviceWindowLoader
        loader = DeviceWindowLoader(ds, runner.device, batch_size=batch, shuffle=True, drop_last=True, seed=CFG.ENV.SEED)
    else:
        loader = torch.utils.data.DataLoader(ds, batch_size=batch, shuffle=True, drop_last=True, pin_memory=True)
    opt = FusedClipAdam([p for p in runner.model.parameters() if p.requires_grad],
                        max_norm=CFG.TRAIN.CLIP_GRAD_PARAM["max_norm"], **CFG.TRAIN.OPTIM.PARAM)
    epoch = 1
    if args.resume:
        epoch = load_checkpoint(args.resume, runner.model, opt)["epoch"] + 1
        print(f"run.py: resumed from {args.resume} (next epoch {epoch})")
    it, t0 = 0, time.perf_counter()
    while it < args.steps:
        for data in loader:
            loss = runner.train_iters(epoch, it, data)
            opt.zero_grad()
            loss.backward()
            opt.step()                                  # global-norm clip + Adam in one pass over the flat gradient buffer
            it += 1
            if it % 5 == 0 or it == args.steps:
                torch.cuda.synchronize()
                print(f"iter {it:4d}  loss {loss.item():.5f}  {it * batch / (time.perf_counter() - t0):.1f} samples/s")
            if it >= args.steps:
                break
    if args.save:
        save_checkpoint(args.save, runner.model, opt, epoch, {"train_loss": float(loss.item())})
        print(f"run.py: checkpoint written to {args.save}")


if __name__ == "__main__":
    main()
