"""`python test/test_inference.py --cfg step/STEP_METR-LA.py --ckpt checkpoints/.../STEP_best_val_MAE.pt --gpus 0` - the
reference's inference CLI (test/test_inference.py:9-35): build the runner from the config, load an easytorch-format
checkpoint, run the test split and print per-horizon and overall MAE / RMSE / MAPE (base_tsf_runner.py:275-318)."""
import importlib
import os
import sys
from argparse import ArgumentParser

sys.path.append(os.path.abspath(__file__ + "/../.."))
import torch  # noqa: E402


def parse_args():
    parser = ArgumentParser(description="Evaluate a trained STEP checkpoint on the B200-native kernels")
    parser.add_argument("-c", "--cfg", default="step/STEP_METR-LA.py", help="training config")
    parser.add_argument("--ckpt", required=True, help="easytorch-format checkpoint")
    parser.add_argument("--gpus", default="0", help="visible gpus")
    parser.add_argument("--synthetic-windows", type=int, default=0, help="evaluate on N synthetic windows instead of the test split")
    return parser.parse_args()


def main():
    args = parse_args()
    os.environ.setdefault("CUDA_VISIBLE_DEVICES", args.gpus)
    from step.step_data import ForecastingDataset
    from step.step_runner.checkpoint import load_checkpoint
    base = os.path.splitext(os.path.basename(args.cfg))[0]
    CFG = importlib.import_module("step." + base).CFG
    runner = CFG.RUNNER(CFG)
    info = load_checkpoint(args.ckpt, runner.model)
    name, seq = CFG.DATASET_NAME, CFG.DATASET_ARGS["seq_len"]
    if args.synthetic_windows > 0:
        ds = ForecastingDataset(mode="test", seq_len=seq, synthetic=True, num_nodes=CFG.MODEL.PARAM["backend_args"]["num_nodes"],
                                length=args.synthetic_windows)
    else:
        d = CFG.TEST.DATA.DIR
        ds = ForecastingDataset(os.path.join(d, "data_in12_out12.pkl"), os.path.join(d, "index_in12_out12.pkl"), "test", seq)
    loader = torch.utils.data.DataLoader(ds, batch_size=CFG.TEST.DATA.BATCH_SIZE, shuffle=False, pin_memory=True)
    rep = runner.test(loader)
    print(f"checkpoint {args.ckpt} (epoch {info['epoch']}, best {info['best_metrics']}) on {len(ds)} windows of {name}")
    for h, m in sorted(rep["horizon"].items()):
        print("Evaluate best model on test data for horizon {0}, Test MAE: {1:.4f}, Test RMSE: {2:.4f}, Test MAPE: {3:.4f}".format(
            h, m["MAE"], m["RMSE"], m["MAPE"]))
    o = rep["overall"]
    print("Result <test>: [test_MAE: {0:.4f}, test_RMSE: {1:.4f}, test_MAPE: {2:.4f}]".format(o["MAE"], o["RMSE"], o["MAPE"]))


if __name__ == "__main__":
    main()
