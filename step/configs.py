"""STEP forecasting configs (stage 2), one factory for the reference's per-dataset files.

Values follow the reference's step/STEP_<NAME>.py (SURVEY.md Appx B): METR-LA / PEMS-BAY / PEMS07 use a 2016-step
(288*7) long history, PEMS03/04/08 a 4032-step one; batch sizes, learning rates and curriculum-learning settings per dataset.
"""
import os

from .easydict_lite import EasyDict
from .step_arch import STEP
from .step_loss import step_loss
from .step_runner import STEPRunner, TSFormerRunner
from .step_data import ForecastingDataset, PretrainingDataset

_NODES = {"METR-LA": 207, "PEMS-BAY": 325, "PEMS03": 358, "PEMS04": 307, "PEMS07": 883, "PEMS08": 170}
_SEQ = {"METR-LA": 288 * 7, "PEMS-BAY": 288 * 7, "PEMS07": 288 * 7, "PEMS03": 288 * 7 * 2, "PEMS04": 288 * 7 * 2, "PEMS08": 288 * 7 * 2}
_BATCH = {"METR-LA": 32, "PEMS-BAY": 32, "PEMS03": 4, "PEMS04": 8, "PEMS07": 4, "PEMS08": 8}
_LR = {"METR-LA": 0.005, "PEMS-BAY": 0.001, "PEMS03": 0.002, "PEMS04": 0.002, "PEMS07": 0.002, "PEMS08": 0.002}
_CL = {"METR-LA": dict(WARM_EPOCHS=0, CL_EPOCHS=6, PREDICTION_LENGTH=12), "PEMS-BAY": dict(WARM_EPOCHS=30, CL_EPOCHS=3, PREDICTION_LENGTH=12)}


def step_config(name: str, gpu_num: int = 1) -> EasyDict:
    CFG = EasyDict()
    CFG.DESCRIPTION = f"STEP({name}) configuration"
    CFG.RUNNER = STEPRunner
    CFG.DATASET_CLS = ForecastingDataset
    CFG.DATASET_NAME = name
    CFG.DATASET_INPUT_LEN = 12
    CFG.DATASET_OUTPUT_LEN = 12
    CFG.DATASET_ARGS = {"seq_len": _SEQ[name]}
    CFG.GPU_NUM = gpu_num
    CFG.ENV = EasyDict(SEED=0, CUDNN=EasyDict(ENABLED=True))
    CFG.MODEL = EasyDict()
    CFG.MODEL.NAME = "STEP"
    CFG.MODEL.ARCH = STEP
    CFG.MODEL.PARAM = {
        "dataset_name": name,
        "pre_trained_tsformer_path": f"tsformer_ckpt/TSFormer_{name}.pt",
        "tsformer_args": {"patch_size": 12, "in_channel": 1, "embed_dim": 96, "num_heads": 4, "mlp_ratio": 4, "dropout": 0.1,
                          "num_token": _SEQ[name] / 12, "mask_ratio": 0.75, "encoder_depth": 4, "decoder_depth": 1,
                          "mode": "forecasting"},
        "backend_args": {"num_nodes": _NODES[name], "support_len": 2, "dropout": 0.3, "gcn_bool": True, "addaptadj": True,
                         "aptinit": None, "in_dim": 2, "out_dim": 12, "residual_channels": 32, "dilation_channels": 32,
                         "skip_channels": 256, "end_channels": 512, "kernel_size": 2, "blocks": 4, "layers": 2},
        "dgl_args": {"dataset_name": name, "k": 10, "input_seq_len": 12, "output_seq_len": 12},
    }
    CFG.MODEL.FORWARD_FEATURES = [0, 1, 2]
    CFG.MODEL.TARGET_FEATURES = [0]
    CFG.MODEL.DDP_FIND_UNUSED_PARAMETERS = True
    CFG.TRAIN = EasyDict()
    CFG.TRAIN.LOSS = step_loss
    CFG.TRAIN.OPTIM = EasyDict(TYPE="Adam", PARAM={"lr": _LR[name], "weight_decay": 1.0e-5, "eps": 1.0e-8})
    CFG.TRAIN.LR_SCHEDULER = EasyDict(TYPE="MultiStepLR", PARAM={"milestones": [1, 18, 36, 54, 72], "gamma": 0.5})
    CFG.TRAIN.CLIP_GRAD_PARAM = {"max_norm": 3.0}
    CFG.TRAIN.NUM_EPOCHS = 100
    CFG.TRAIN.CKPT_SAVE_DIR = os.path.join("checkpoints", "STEP_100")
    CFG.TRAIN.NULL_VAL = 0.0
    CFG.TRAIN.DATA = EasyDict(DIR="datasets/" + name, BATCH_SIZE=_BATCH[name], PREFETCH=False, SHUFFLE=True, NUM_WORKERS=2,
                              PIN_MEMORY=True)
    if name in _CL:
        CFG.TRAIN.CL = EasyDict(_CL[name])
    for split in ("VAL", "TEST"):
        CFG[split] = EasyDict(INTERVAL=1, DATA=EasyDict(DIR="datasets/" + name, BATCH_SIZE=_BATCH[name], PREFETCH=False,
                                                        SHUFFLE=False, NUM_WORKERS=2, PIN_MEMORY=True))
    return CFG


# ---- stage 1: TSFormer pre-training configs (reference step/TSFormer_<NAME>.py) ------------------------------------
_TS_BATCH = {"METR-LA": 8, "PEMS-BAY": 16, "PEMS03": 3, "PEMS04": 6, "PEMS07": 3, "PEMS08": 6}
_TS_SEQ = dict(_SEQ, PEMS03=288 * 7)     # the reference pre-trains PEMS03 on one week although STEP_PEMS03 reads two
_TS_LR = {"METR-LA": 0.0005, "PEMS-BAY": 0.001, "PEMS03": 0.001, "PEMS04": 0.001, "PEMS07": 0.001, "PEMS08": 0.001}


def tsformer_config(name: str, gpu_num: int = 1) -> EasyDict:
    """Masked-patch pre-training of TSFormer: history of _TS_SEQ[name] steps, channel 0 only, 75 % of the 12-step patches
    masked, objective masked MAE (null value 0) between the reconstructed and the true masked patches."""
    from .step_arch import TSFormer
    from .step_runner.metrics import masked_mae
    CFG = EasyDict()
    CFG.DESCRIPTION = f"TSFormer({name}) configuration"
    CFG.RUNNER = TSFormerRunner
    CFG.DATASET_CLS = PretrainingDataset
    CFG.DATASET_NAME = name
    CFG.DATASET_INPUT_LEN = _TS_SEQ[name]
    CFG.DATASET_OUTPUT_LEN = 12
    CFG.DATASET_ARGS = {}
    CFG.GPU_NUM = gpu_num
    CFG.ENV = EasyDict(SEED=0, CUDNN=EasyDict(ENABLED=True))
    CFG.MODEL = EasyDict()
    CFG.MODEL.NAME = "TSFormer"
    CFG.MODEL.ARCH = TSFormer
    CFG.MODEL.PARAM = {"patch_size": 12, "in_channel": 1, "embed_dim": 96, "num_heads": 4, "mlp_ratio": 4, "dropout": 0.1,
                       "num_token": _TS_SEQ[name] / 12, "mask_ratio": 0.75, "encoder_depth": 4, "decoder_depth": 1,
                       "mode": "pre-train"}
    CFG.MODEL.FORWARD_FEATURES = [0]
    CFG.MODEL.TARGET_FEATURES = [0]
    CFG.TRAIN = EasyDict()
    CFG.TRAIN.LOSS = masked_mae
    CFG.TRAIN.OPTIM = EasyDict(TYPE="Adam", PARAM={"lr": _TS_LR[name], "weight_decay": 0, "eps": 1.0e-8, "betas": (0.9, 0.95)})
    CFG.TRAIN.LR_SCHEDULER = EasyDict(TYPE="MultiStepLR", PARAM={"milestones": [50], "gamma": 0.5})
    CFG.TRAIN.CLIP_GRAD_PARAM = {"max_norm": 5.0}
    CFG.TRAIN.NUM_EPOCHS = 100
    CFG.TRAIN.CKPT_SAVE_DIR = os.path.join("checkpoints", "TSFormer_100")
    CFG.TRAIN.NULL_VAL = 0.0
    CFG.TRAIN.DATA = EasyDict(DIR="datasets/" + name, BATCH_SIZE=_TS_BATCH[name], PREFETCH=False, SHUFFLE=True, NUM_WORKERS=2,
                              PIN_MEMORY=True)
    for split in ("VAL", "TEST"):
        CFG[split] = EasyDict(INTERVAL=1, DATA=EasyDict(DIR="datasets/" + name, BATCH_SIZE=_TS_BATCH[name], PREFETCH=False,
                                                        SHUFFLE=False, NUM_WORKERS=2, PIN_MEMORY=True))
    return CFG
