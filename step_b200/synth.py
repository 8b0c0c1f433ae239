"""Deterministic synthetic parameters and inputs for STEP (seeded torch generators; no file or network access).

Used by ``bench.py`` (the product arm measures on these tensors: there is no network for datasets or checkpoints), by the
tests, and re-exported by ``oracle/step_oracle.py`` for the golden-fixture scripts.  Pure data generation: shapes follow the
reference's modules (``graphwavenet/model.py:51-119``, ``discrete_graph_learning.py:51-78``, SURVEY.md Appx C for the 72
TSFormer keys), values are N(0, fan-in-scaled) / N(0,1) draws - nothing here computes any part of the STEP path.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

Tensor = torch.Tensor
SD = Dict[str, Tensor]

# --------------------------------------------------------------------------- #
# dataset tables  (step/step_arch/discrete_graph_learning.py:55-56,61,73)
# --------------------------------------------------------------------------- #
NUM_NODES = {"METR-LA": 207, "PEMS04": 307, "PEMS03": 358, "PEMS-BAY": 325, "PEMS07": 883, "PEMS08": 170}
TRAIN_LENGTH = {"METR-LA": 23990, "PEMS04": 13599, "PEMS03": 15303, "PEMS07": 16513, "PEMS-BAY": 36482, "PEMS08": 14284}
DIM_FC = {"METR-LA": 383552, "PEMS04": 217296, "PEMS03": 244560, "PEMS07": 263920, "PEMS-BAY": 583424, "PEMS08": 228256}
DIM_FC_MEAN = {"METR-LA": 16128, "PEMS-BAY": 16128, "PEMS03": 32256, "PEMS04": 32256, "PEMS07": 16128, "PEMS08": 32256}


# --------------------------------------------------------------------------- #
# deterministic synthetic parameters / inputs shared by tests, golden script and bench
# --------------------------------------------------------------------------- #
def gwnet_param_shapes(num_nodes: int):
    shp = {"start_conv.weight": (32, 2, 1, 1), "start_conv.bias": (32,),
           "nodevec1": (num_nodes, 10), "nodevec2": (10, num_nodes),
           "fc_his.0.weight": (512, 96), "fc_his.0.bias": (512,), "fc_his.2.weight": (256, 512), "fc_his.2.bias": (256,),
           "end_conv_1.weight": (512, 256, 1, 1), "end_conv_1.bias": (512,),
           "end_conv_2.weight": (12, 512, 1, 1), "end_conv_2.bias": (12,)}
    for i in range(8):
        shp[f"filter_convs.{i}.weight"] = (32, 32, 1, 2); shp[f"filter_convs.{i}.bias"] = (32,)
        shp[f"gate_convs.{i}.weight"] = (32, 32, 1, 2); shp[f"gate_convs.{i}.bias"] = (32,)
        shp[f"residual_convs.{i}.weight"] = (32, 32, 1, 1); shp[f"residual_convs.{i}.bias"] = (32,)
        shp[f"skip_convs.{i}.weight"] = (256, 32, 1, 1); shp[f"skip_convs.{i}.bias"] = (256,)
        shp[f"bn.{i}.weight"] = (32,); shp[f"bn.{i}.bias"] = (32,)
        shp[f"gconv.{i}.mlp.mlp.weight"] = (32, 224, 1, 1); shp[f"gconv.{i}.mlp.mlp.bias"] = (32,)
    return shp


def dgl_param_shapes(dataset: str):
    return {"conv1.weight": (8, 1, 10), "conv1.bias": (8,), "conv2.weight": (16, 8, 10), "conv2.bias": (16,),
            "fc.weight": (100, DIM_FC[dataset]), "fc.bias": (100,),
            "bn1.weight": (8,), "bn1.bias": (8,), "bn2.weight": (16,), "bn2.bias": (16,),
            "bn3.weight": (100,), "bn3.bias": (100,),
            "fc_mean.weight": (100, DIM_FC_MEAN[dataset]), "fc_mean.bias": (100,),
            "fc_cat.weight": (2, 100), "fc_cat.bias": (2,), "fc_out.weight": (100, 200), "fc_out.bias": (100,)}


def _fan_in(shape) -> int:
    n = 1
    for s in shape[1:]:
        n *= s
    return max(n, 1)


def synthetic_trainable_params(dataset: str, seed: int = 0) -> SD:
    """Deterministic GWNet/DGL parameters (keys = reference state-dict keys).  Weights ~
    U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (torch's default Linear/Conv bound), BN weight/bias
    perturbed around (1, 0) so that BN parameters matter in parity tests, nodevec ~ N(0,1)
    (graphwavenet/model.py:83-84).  Each tensor has its own generator so the values do not
    depend on construction order."""
    out: SD = {}
    n = NUM_NODES[dataset]
    groups = (("backend.", gwnet_param_shapes(n)), ("discrete_graph_learning.", dgl_param_shapes(dataset)))
    idx = 0
    for prefix, shapes in groups:
        for name, shape in shapes.items():
            g = torch.Generator().manual_seed(seed * 100003 + idx)
            idx += 1
            if name.startswith("nodevec"):
                t = torch.randn(shape, generator=g)
            elif ".bn" in "." + name or name.startswith("bn"):
                t = (1.0 if name.endswith("weight") else 0.0) + 0.1 * (torch.rand(shape, generator=g) * 2 - 1)
            else:
                fan = _fan_in(shape) if name.endswith("weight") else None
                if fan is None:
                    wshape = shapes[name[:-4] + "weight"]
                    fan = _fan_in(wshape)
                bound = 1.0 / math.sqrt(fan)
                t = (torch.rand(shape, generator=g) * 2 - 1) * bound
            out[prefix + name] = t
    return out


def bn_buffers(dataset: str) -> SD:
    """Fresh BatchNorm buffers (running_mean 0, running_var 1, num_batches_tracked 0)."""
    out: SD = {}
    for i in range(8):
        out[f"backend.bn.{i}.running_mean"] = torch.zeros(32)
        out[f"backend.bn.{i}.running_var"] = torch.ones(32)
        out[f"backend.bn.{i}.num_batches_tracked"] = torch.zeros((), dtype=torch.long)
    for name, c in (("bn1", 8), ("bn2", 16), ("bn3", 100)):
        out[f"discrete_graph_learning.{name}.running_mean"] = torch.zeros(c)
        out[f"discrete_graph_learning.{name}.running_var"] = torch.ones(c)
        out[f"discrete_graph_learning.{name}.num_batches_tracked"] = torch.zeros((), dtype=torch.long)
    return out


def synthetic_tsformer_params(seed: int = 0, embed: int = 96, patch: int = 12, depth: int = 4, dec_depth: int = 1) -> SD:
    """Random TSFormer weights with the 72 checkpoint keys (SURVEY Appx C) and magnitudes
    close to the shipped checkpoints' (so that softmax sharpness is realistic)."""
    g = torch.Generator().manual_seed(seed + 7919)
    r = lambda *s, sc=0.09: torch.randn(*s, generator=g) * sc
    sd: SD = {"mask_token": r(1, 1, 1, embed, sc=0.05),
              "encoder_norm.weight": 0.6 + r(embed, sc=0.1), "encoder_norm.bias": r(embed, sc=0.1),
              "decoder_norm.weight": 0.7 + r(embed, sc=0.1), "decoder_norm.bias": r(embed, sc=0.1),
              "patch_embedding.input_embedding.weight": r(embed, 1, patch, 1, sc=0.12),
              "patch_embedding.input_embedding.bias": r(embed, sc=0.18),
              "positional_encoding.position_embedding": r(1000, embed, sc=0.03),
              "enc_2_dec_emb.weight": r(embed, embed), "enc_2_dec_emb.bias": r(embed, sc=0.1),
              "output_layer.weight": r(patch, embed, sc=0.05), "output_layer.bias": r(patch, sc=0.05)}
    for stack, n in (("encoder", depth), ("decoder", dec_depth)):
        for i in range(n):
            p = f"{stack}.transformer_encoder.layers.{i}."
            sd[p + "self_attn.in_proj_weight"] = r(3 * embed, embed, sc=0.11)
            sd[p + "self_attn.in_proj_bias"] = r(3 * embed, sc=0.05)
            sd[p + "self_attn.out_proj.weight"] = r(embed, embed, sc=0.1)
            sd[p + "self_attn.out_proj.bias"] = r(embed, sc=0.04)
            sd[p + "linear1.weight"] = r(4 * embed, embed, sc=0.13)
            sd[p + "linear1.bias"] = r(4 * embed, sc=0.2)
            sd[p + "linear2.weight"] = r(embed, 4 * embed, sc=0.11)
            sd[p + "linear2.bias"] = r(embed, sc=0.12)
            sd[p + "norm1.weight"] = 0.85 + r(embed, sc=0.1); sd[p + "norm1.bias"] = r(embed, sc=0.15)
            sd[p + "norm2.weight"] = 0.82 + r(embed, sc=0.1); sd[p + "norm2.bias"] = r(embed, sc=0.08)
    return sd


def synthetic_batch(dataset: str, batch: int, patches: int, seed: int = 0):
    """SURVEY section 8(d) inputs: history/future ~ N(0,1) [B,12,N,3], long_history ~ N(0,1)
    [B,P*12,N,3], Gumbel uniforms U(0,1) [B,N*N,2]."""
    n = NUM_NODES[dataset]
    g = torch.Generator().manual_seed(seed + 1)
    history = torch.randn(batch, 12, n, 3, generator=g)
    long_history = torch.randn(batch, patches * 12, n, 3, generator=g)
    future = torch.randn(batch, 12, n, 3, generator=g)
    uniform = torch.rand(batch, n * n, 2, generator=g)
    return history, long_history, future, uniform


def synthetic_node_feats(dataset: str, seed: int = 0) -> Tensor:
    """[train_len, N] standard normal 'training series' (discrete_graph_learning.py:57)."""
    g = torch.Generator().manual_seed(seed + 2)
    return torch.randn(TRAIN_LENGTH[dataset], NUM_NODES[dataset], generator=g)
