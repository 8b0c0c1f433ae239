import os
import pickle
import sys

import pytest
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


TS_ARGS = dict(patch_size=12, in_channel=1, embed_dim=96, num_heads=4, mlp_ratio=4, dropout=0.1, num_token=168,
               mask_ratio=0.75, encoder_depth=4, decoder_depth=1, mode="forecasting")


def gw_args(n):
    return dict(num_nodes=n, support_len=2, dropout=0.3, gcn_bool=True, addaptadj=True, aptinit=None, in_dim=2, out_dim=12,
                residual_channels=32, dilation_channels=32, skip_channels=256, end_channels=512, kernel_size=2, blocks=4,
                layers=2)


def build_step_model(tmp_path, dataset, seed=0, real_ckpt=False):
    """Construct our STEP module the way the runner does (cfg.MODEL.ARCH(**cfg.MODEL.PARAM)) with the
    deterministic synthetic parameters the golden fixtures were generated with."""
    from oracle import step_oracle as O
    from step.step_arch import STEP
    n = O.NUM_NODES[dataset]
    node_feats = O.synthetic_node_feats(dataset, seed)
    d = tmp_path / "datasets" / dataset
    d.mkdir(parents=True, exist_ok=True)
    with open(d / "data_in12_out12.pkl", "wb") as f:
        pickle.dump({"processed_data": node_feats.unsqueeze(-1).numpy()}, f)
    if real_ckpt:
        ts_sd = torch.load(os.path.join(GOLDEN, f"tsformer_{dataset}_state.pt"))
    else:
        ts_sd = O.synthetic_tsformer_params(seed)
    torch.save({"model_state_dict": ts_sd}, tmp_path / "ts.pt")
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        model = STEP(dataset, str(tmp_path / "ts.pt"), dict(TS_ARGS), gw_args(n),
                     dict(dataset_name=dataset, k=10, input_seq_len=12, output_seq_len=12))
    finally:
        os.chdir(cwd)
    full = dict(O.synthetic_trainable_params(dataset, seed))
    full.update(O.bn_buffers(dataset))
    full.update({"tsformer." + k: v for k, v in ts_sd.items()})
    model.load_state_dict(full, strict=True)
    return model, full, node_feats
