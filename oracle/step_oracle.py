"""CPU oracle for the STEP forward/backward hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``step_b200/`` or ``step/`` may import
this file; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker / the
CPU comparator, never as the product path.

It is a *functional restatement* (plain torch CPU ops on a flat ``{name: tensor}``
state dict that uses the reference's state-dict keys) of what the reference's
``step/step_arch`` modules compute.  Every function cites the reference lines it
restates (paths relative to the reference repo root).

Parity status: the reference ships no tests or golden vectors, so
the oracle is pinned against *the reference itself*, imported in the build
container by ``tests/golden/make_golden.py``; the committed fixtures under
``tests/golden/*.pt`` are the reference's outputs and the CPU suite re-checks the
oracle against them (``tests/test_oracle_golden.py``).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

# --------------------------------------------------------------------------- #
# dataset tables and deterministic synthetic parameters / inputs: defined once in step_b200/synth.py (pure data
# generation shared with bench.py's product arm, which must not import this oracle) and re-exported here
# --------------------------------------------------------------------------- #
from step_b200.synth import (DIM_FC, DIM_FC_MEAN, NUM_NODES, TRAIN_LENGTH, bn_buffers, dgl_param_shapes,  # noqa: E402,F401
                             gwnet_param_shapes, synthetic_batch, synthetic_node_feats, synthetic_trainable_params,
                             synthetic_tsformer_params)

BN_EPS = 1e-5
LN_EPS = 1e-5


# --------------------------------------------------------------------------- #
# TSFormer encoder, forecasting mode
# --------------------------------------------------------------------------- #
def _layer_norm(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + LN_EPS) * w + b


def _drop(x: Tensor, p: float) -> Tensor:
    return F.dropout(x, p, training=True) if p > 0.0 else x


USE_SDPA = False      # bench.py's GPU-eager comparator sets this: torch's nn.TransformerEncoderLayer (what the reference
                      # instantiates, transformer_layers.py:10-11) reaches F.scaled_dot_product_attention in train mode


def encoder_layer(sd: SD, pre: str, z: Tensor, heads: int, p_drop: float = 0.0) -> Tensor:
    """One post-norm ``nn.TransformerEncoderLayer(d, heads, 4d, dropout)`` with ReLU
    (step/step_arch/tsformer/transformer_layers.py:10-11; formula of the Python
    "slow path" of torch's layer, SURVEY Appx A.2).  ``z``: [S, P, d] (batch-first
    here; the reference is sequence-first, transformer_layers.py:16-18 - the math
    is per sequence so the order of the first two axes is immaterial)."""
    S, P, d = z.shape
    hd = d // heads
    qkv = z @ sd[pre + "self_attn.in_proj_weight"].t() + sd[pre + "self_attn.in_proj_bias"]
    q, k, v = qkv.split(d, dim=-1)
    q = q.view(S, P, heads, hd).transpose(1, 2)
    k = k.view(S, P, heads, hd).transpose(1, 2)
    v = v.view(S, P, heads, hd).transpose(1, 2)
    if USE_SDPA:
        o = F.scaled_dot_product_attention(q, k, v, dropout_p=p_drop).transpose(1, 2).reshape(S, P, d)
    else:
        att = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
        att = _drop(att, p_drop)
        o = (att @ v).transpose(1, 2).reshape(S, P, d)
    o = o @ sd[pre + "self_attn.out_proj.weight"].t() + sd[pre + "self_attn.out_proj.bias"]
    z = _layer_norm(z + _drop(o, p_drop), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"])
    f = torch.relu(z @ sd[pre + "linear1.weight"].t() + sd[pre + "linear1.bias"])
    f = _drop(f, p_drop) @ sd[pre + "linear2.weight"].t() + sd[pre + "linear2.bias"]
    z = _layer_norm(z + _drop(f, p_drop), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"])
    return z


def tsformer_tokens(sd: SD, series: Tensor, pre: str = "") -> Tensor:
    """Patch embedding + learned positional embedding.
    ``series``: [S, P*L] -> tokens [S, P, d].
    step/step_arch/tsformer/patch.py:31-42 (Conv2d(1,d,(L,1),stride (L,1)) == a
    [L]->[d] linear map on each non-overlapping patch) and
    positional_encoding.py:24-35 (``x + pos[:P]``)."""
    w = sd[pre + "patch_embedding.input_embedding.weight"]          # [d,1,L,1]
    b = sd[pre + "patch_embedding.input_embedding.bias"]
    d, _, L, _ = w.shape
    S, T = series.shape
    P = T // L
    tok = series.view(S, P, L) @ w.view(d, L).t() + b
    return tok + sd[pre + "positional_encoding.position_embedding"][:P]


def tsformer_encode(sd: SD, long_history: Tensor, pre: str = "", heads: int = 4, depth: int = 4,
                    p_drop: float = 0.0) -> Tensor:
    """TSFormer.forward(mode="forecasting"): step/step_arch/tsformer/tsformer.py:179,189-191
    -> encoding(mask=False) :86-105.  ``long_history``: [B, P*L, N, 1] -> [B, N, P, d]."""
    B, T, N, _ = long_history.shape
    series = long_history[..., 0].permute(0, 2, 1).reshape(B * N, T)
    z = tsformer_tokens(sd, series, pre)
    z = _drop(z, p_drop)                                              # positional_encoding.py:32
    d = z.shape[-1]
    z = z * math.sqrt(d)                                              # transformer_layers.py:15
    for i in range(depth):
        z = encoder_layer(sd, f"{pre}encoder.transformer_encoder.layers.{i}.", z, heads, p_drop)
    z = _layer_norm(z, sd[pre + "encoder_norm.weight"], sd[pre + "encoder_norm.bias"])   # tsformer.py:103
    return z.view(B, N, -1, d)


def tsformer_pretrain(sd: SD, history: Tensor, unmasked: list, masked: list, pre: str = "", heads: int = 4,
                      depth: int = 4, dec_depth: int = 1, p_drop: float = 0.0):
    """TSFormer.forward(mode="pre-train") with the mask draw given (the reference draws it with Python's
    ``random``, tsformer/mask.py:15-24): step/step_arch/tsformer/tsformer.py:179-188 ->
    encoding :86-105 (patchify, positional embedding, keep the unmasked tokens, encoder, encoder_norm),
    decoding :107-136 (enc_2_dec_emb, mask tokens + positional embedding of the masked positions, decoder,
    decoder_norm, output_layer), get_reconstructed_masked_tokens :138-160.
    ``history``: [B, P*L, N, C] -> (reconstruction [B, r*P*L, N], label [B, r*P*L, N])."""
    B, T, N, _ = history.shape
    series = history[..., 0].permute(0, 2, 1).reshape(B * N, T)
    z = _drop(tsformer_tokens(sd, series, pre), p_drop)                       # [S, P, d]
    S, P, d = z.shape
    L = T // P
    z = z[:, unmasked, :] * math.sqrt(d)
    for i in range(depth):
        z = encoder_layer(sd, f"{pre}encoder.transformer_encoder.layers.{i}.", z, heads, p_drop)
    z = _layer_norm(z, sd[pre + "encoder_norm.weight"], sd[pre + "encoder_norm.bias"])
    z = z @ sd[pre + "enc_2_dec_emb.weight"].t() + sd[pre + "enc_2_dec_emb.bias"]
    m = sd[pre + "mask_token"].view(1, 1, d) + sd[pre + "positional_encoding.position_embedding"][masked].unsqueeze(0)
    m = _drop(m.expand(S, len(masked), d), p_drop)
    z = torch.cat([z, m], dim=1) * math.sqrt(d)
    for i in range(dec_depth):
        z = encoder_layer(sd, f"{pre}decoder.transformer_encoder.layers.{i}.", z, heads, p_drop)
    z = _layer_norm(z, sd[pre + "decoder_norm.weight"], sd[pre + "decoder_norm.bias"])
    rec = (z @ sd[pre + "output_layer.weight"].t() + sd[pre + "output_layer.bias"]).view(B, N, P, L)
    rec = rec[:, :, len(unmasked):, :].reshape(B, N, -1).transpose(1, 2)
    label = series.view(B, N, P, L)[:, :, masked, :].reshape(B, N, -1).transpose(1, 2)
    return rec, label


# --------------------------------------------------------------------------- #
# Discrete graph learning
# --------------------------------------------------------------------------- #
def _bn_train(x: Tensor, w: Tensor, b: Tensor, dims) -> Tensor:
    mu = x.mean(dims, keepdim=True)
    var = ((x - mu) ** 2).mean(dims, keepdim=True)
    shape = [1] * x.dim()
    shape[1] = -1
    return (x - mu) / torch.sqrt(var + BN_EPS) * w.view(shape) + b.view(shape)


def _bn_eval(x: Tensor, w: Tensor, b: Tensor, rm: Tensor, rv: Tensor) -> Tensor:
    shape = [1] * x.dim()
    shape[1] = -1
    return (x - rm.view(shape)) / torch.sqrt(rv.view(shape) + BN_EPS) * w.view(shape) + b.view(shape)


def _bn(sd: SD, pre: str, x: Tensor, train: bool) -> Tensor:
    if train:
        return _bn_train(x, sd[pre + "weight"], sd[pre + "bias"], [i for i in range(x.dim()) if i != 1])
    return _bn_eval(x, sd[pre + "weight"], sd[pre + "bias"], sd[pre + "running_mean"], sd[pre + "running_var"])


def dgl_trunk(sd: SD, node_feats: Tensor, pre: str = "discrete_graph_learning.", train: bool = True) -> Tensor:
    """Batch-invariant "global feature" CNN: discrete_graph_learning.py:131-135.
    ``node_feats``: [train_len, N] -> [N, 100]."""
    x = node_feats.t().unsqueeze(1)                                   # [N,1,L]
    x = _bn(sd, pre + "bn1.", torch.relu(F.conv1d(x, sd[pre + "conv1.weight"], sd[pre + "conv1.bias"])), train)
    x = _bn(sd, pre + "bn2.", torch.relu(F.conv1d(x, sd[pre + "conv2.weight"], sd[pre + "conv2.bias"])), train)
    x = x.reshape(x.shape[0], -1)
    x = torch.relu(x @ sd[pre + "fc.weight"].t() + sd[pre + "fc.bias"])
    return _bn(sd, pre + "bn3.", x, train)


def edge_logits(sd: SD, feat: Tensor, pre: str = "discrete_graph_learning.") -> Tensor:
    """Edge MLP, discrete_graph_learning.py:148-153.  Edge e = i*N + j has receiver
    feat[i] and sender feat[j] (rel_rec[e] = onehot(e // N), rel_send[e] = onehot(e % N),
    :81-89) and the concat order is [senders, receivers] (:150).  Returns [N*N, 2]."""
    N = feat.shape[0]
    w_out, b_out = sd[pre + "fc_out.weight"], sd[pre + "fc_out.bias"]          # [100,200]
    half = w_out.shape[1] // 2
    u = feat @ w_out[:, :half].t()       # sender part, indexed by j
    v = feat @ w_out[:, half:].t()       # receiver part, indexed by i
    h = torch.relu(v[:, None, :] + u[None, :, :] + b_out)                      # [i, j, 100]
    return (h @ sd[pre + "fc_cat.weight"].t() + sd[pre + "fc_cat.bias"]).reshape(N * N, 2)


def gumbel_hard_sample(logits: Tensor, uniform: Tensor, tau: float = 0.5, eps: float = 1e-10) -> Tensor:
    """discrete_graph_learning.py:11-45 with hard=True: g = -log(-log(U+eps)+eps),
    y = softmax((logits+g)/tau), one-hot argmax with straight-through gradient.
    logits [..., 2] broadcastable to uniform [B, N*N, 2]; returns y [B, N*N, 2]."""
    g = -torch.log(-torch.log(uniform + eps) + eps)
    y = torch.softmax((logits + g) / tau, dim=-1)
    idx = y.detach().argmax(-1, keepdim=True)
    hard = torch.zeros_like(y).scatter_(-1, idx, 1.0)
    return (hard - y.detach()) + y


def cosine_similarity_gram(x: Tensor) -> Tensor:
    """step/step_arch/similarity.py:6-16 with y = x.  x [B,N,D] -> [B,N,N]."""
    n = x.norm(dim=2) + 1e-7
    return (x @ x.transpose(1, 2)) / (n[:, :, None] * n[:, None, :])


def knn_prior(hidden: Tensor, k_total: int) -> Tensor:
    """get_k_nn_neighbor (discrete_graph_learning.py:91-111) + diagonal removal (:165-166).
    hidden [B,N,P,d] -> adj_knn [B,N,N] in {0,1}; *global* top-k_total of the N*N entries."""
    B, N = hidden.shape[:2]
    sim = cosine_similarity_gram(hidden.reshape(B, N, -1)).reshape(B, N * N)
    vals, idx = torch.topk(sim, k_total, dim=-1)
    res = torch.zeros_like(sim).scatter_(-1, idx, vals)
    adj = (res != 0).to(sim.dtype).view(B, N, N)
    eye = torch.eye(N, dtype=torch.bool, device=sim.device)
    return adj.masked_fill(eye, 0.0).detach()


def discrete_graph_learning(sd: SD, long_history: Tensor, node_feats: Tensor, k: int, uniform: Tensor,
                            train: bool = True, pre: str = "discrete_graph_learning.",
                            ts_pre: str = "tsformer.", ts_drop: float = 0.0):
    """DiscreteGraphLearning.forward, discrete_graph_learning.py:113-168.
    Returns (bernoulli_unnorm [B,N*N,2], hidden [B,N,P,d], adj_knn [B,N,N], sampled_adj [B,N,N])."""
    B, _, N, _ = long_history.shape
    feat = dgl_trunk(sd, node_feats, pre, train)
    with torch.no_grad():
        hidden = tsformer_encode(sd, long_history[..., [0]], ts_pre, p_drop=ts_drop)
    logits = edge_logits(sd, feat, pre)                               # identical for every sample (Appx D.1)
    bern = logits.unsqueeze(0).expand(B, N * N, 2)
    y = gumbel_hard_sample(bern, uniform)
    eye = torch.eye(N, dtype=torch.bool, device=y.device)
    sampled = y[..., 0].reshape(B, N, N).masked_fill(eye, 0.0)
    adj_knn = knn_prior(hidden, k * N)
    return bern, hidden, adj_knn, sampled


# --------------------------------------------------------------------------- #
# Graph WaveNet backbone
# --------------------------------------------------------------------------- #
def random_walk(adj: Tensor) -> Tensor:
    """GraphWaveNet._calculate_random_walk_matrix, graphwavenet/model.py:121-130: D^-1 (A + I)."""
    N = adj.shape[1]
    a = adj + torch.eye(N, dtype=adj.dtype, device=adj.device)
    d = a.sum(2)
    dinv = torch.where(d == 0, torch.zeros_like(d), 1.0 / d)
    return dinv.unsqueeze(-1) * a


def _mix_nodes(x: Tensor, a: Tensor) -> Tensor:
    """nconv, graphwavenet/model.py:10-16: out[n,c,w,l] = sum_v x[n,c,v,l] a[(n,)v,w]."""
    if a.dim() == 3:
        return torch.einsum("ncvl,nvw->ncwl", x, a)
    return torch.einsum("ncvl,vw->ncwl", x, a)


def gwnet_forward(sd: SD, history: Tensor, hidden_last: Tensor, sampled_adj: Tensor, pre: str = "backend.",
                  train: bool = True, p_drop: float = 0.0, n_layers: int = 8, taps=None) -> Tensor:
    """GraphWaveNet.forward, graphwavenet/model.py:132-224.
    history [B,12,N,3], hidden_last [B,N,96], sampled_adj [B,N,N] -> [B,N,12].
    ``taps`` (optional dict) receives per-layer intermediates for kernel-level tests."""
    x = F.pad(history.transpose(1, 3), (1, 0))[:, :2]                 # [B,2,N,13]   :145-149
    x = F.conv2d(x, sd[pre + "start_conv.weight"], sd[pre + "start_conv.bias"])
    supports = [random_walk(sampled_adj), random_walk(sampled_adj.transpose(-1, -2)),
                torch.softmax(torch.relu(sd[pre + "nodevec1"] @ sd[pre + "nodevec2"]), dim=1)]   # :160-166
    skip = None
    for i in range(n_layers):
        dil = 1 if i % 2 == 0 else 2
        res = x
        f = torch.tanh(F.conv2d(res, sd[f"{pre}filter_convs.{i}.weight"], sd[f"{pre}filter_convs.{i}.bias"], dilation=(1, dil)))
        g = torch.sigmoid(F.conv2d(res, sd[f"{pre}gate_convs.{i}.weight"], sd[f"{pre}gate_convs.{i}.bias"], dilation=(1, dil)))
        x = f * g
        s = F.conv2d(x, sd[f"{pre}skip_convs.{i}.weight"], sd[f"{pre}skip_convs.{i}.bias"])
        skip = s if skip is None else s + skip[..., -s.shape[3]:]
        outs = [x]
        for a in supports:
            x1 = _mix_nodes(x, a)
            x2 = _mix_nodes(x1, a)
            outs += [x1, x2]
        h = F.conv2d(torch.cat(outs, dim=1), sd[f"{pre}gconv.{i}.mlp.mlp.weight"], sd[f"{pre}gconv.{i}.mlp.mlp.bias"])
        h = _drop(h, p_drop) if train else h
        x = h + res[..., -h.shape[3]:]
        if taps is not None:
            taps[f"z{i}"] = x
        x = _bn(sd, f"{pre}bn.{i}.", x, train)
    hs = torch.relu(hidden_last @ sd[pre + "fc_his.0.weight"].t() + sd[pre + "fc_his.0.bias"])
    hs = torch.relu(hs @ sd[pre + "fc_his.2.weight"].t() + sd[pre + "fc_his.2.bias"])
    skip = skip + hs.transpose(1, 2).unsqueeze(-1)
    if taps is not None:
        taps["skip"] = skip
    x = torch.relu(skip)
    x = torch.relu(F.conv2d(x, sd[pre + "end_conv_1.weight"], sd[pre + "end_conv_1.bias"]))
    x = F.conv2d(x, sd[pre + "end_conv_2.weight"], sd[pre + "end_conv_2.bias"])
    return x.squeeze(-1).transpose(1, 2)


# --------------------------------------------------------------------------- #
# whole model + loss
# --------------------------------------------------------------------------- #
def step_forward(sd: SD, history: Tensor, long_history: Tensor, node_feats: Tensor, uniform: Tensor,
                 epoch: Optional[int] = 1, k: int = 10, train: bool = True, gw_drop: float = 0.0,
                 ts_drop: float = 0.0):
    """STEP.forward, step/step_arch/step.py:37-72.
    Returns (y_hat [B,12,N,1], theta [B,N,N], adj_knn [B,N,N], gsl_coefficient)."""
    B, _, N, _ = history.shape
    bern, hidden, adj_knn, sampled = discrete_graph_learning(sd, long_history, node_feats, k, uniform, train,
                                                             ts_drop=ts_drop)
    y = gwnet_forward(sd, history, hidden[:, :, -1, :], sampled, train=train, p_drop=gw_drop).transpose(1, 2)
    coeff = 1 / (int(epoch / 6) + 1) if epoch is not None else 0
    theta = torch.softmax(bern, -1)[..., 0].clone().reshape(B, N, N)
    return y.unsqueeze(-1), theta, adj_knn, coeff


def masked_mae(preds: Tensor, labels: Tensor, null_val: float = float("nan")) -> Tensor:
    """basicts/metrics/mae.py:5-28."""
    if math.isnan(null_val):
        mask = ~torch.isnan(labels)
    else:
        mask = (labels - null_val).abs() > 5e-5
    mask = mask.float()
    mask = mask / mask.mean()
    mask = torch.where(torch.isnan(mask), torch.zeros_like(mask), mask)
    loss = (preds - labels).abs() * mask
    loss = torch.where(torch.isnan(loss), torch.zeros_like(loss), loss)
    return loss.mean()


def step_loss(prediction: Tensor, real_value: Tensor, theta: Tensor, priori_adj: Tensor, gsl_coefficient: float,
              null_val: float = float("nan")) -> Tensor:
    """step/step_loss/step_loss.py:5-16 (nn.BCELoss clamps log at -100)."""
    B, N, _ = theta.shape
    th = theta.reshape(B, N * N)
    tr = priori_adj.reshape(B, N * N)
    bce = -(tr * torch.log(th).clamp_min(-100.0) + (1 - tr) * torch.log(1 - th).clamp_min(-100.0)).mean()
    return masked_mae(prediction, real_value, null_val) + bce * gsl_coefficient


def train_step(sd: SD, history: Tensor, long_history: Tensor, future: Tensor, node_feats: Tensor, uniform: Tensor,
               epoch: int = 1, mean: float = 0.0, std: float = 1.0, null_val: float = 0.0, gw_drop: float = 0.0,
               ts_drop: float = 0.0, adj_knn_override: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """One fwd+loss as the runner performs it (basicts/runners/base_tsf_runner.py:237-250,
    step/step_runner/step_runner.py:57-75): target feature 0, re_standard_transform, step_loss."""
    y_hat, theta, adj_knn, coeff = step_forward(sd, history, long_history, node_feats, uniform, epoch,
                                                gw_drop=gw_drop, ts_drop=ts_drop)
    if adj_knn_override is not None:      # isolate top-k tie-breaking from everything else in tests
        adj_knn = adj_knn_override
    pred = y_hat[..., [0]] * std + mean
    real = future[..., [0]] * std + mean
    return step_loss(pred, real, theta, adj_knn, coeff, null_val), y_hat
