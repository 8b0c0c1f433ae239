"""Discrete graph learning on the B200-native kernels (drop-in for the reference module).

Reference: ``step/step_arch/discrete_graph_learning.py:48-168``.  Same constructor, ``forward``
signature/returns and state-dict keys.  Differences in *how* (not what):
  * the N^2 x N one-hot matmuls (rel_rec / rel_send, :88-89,148-149) are an index gather folded into
    the edge-logit kernel, evaluated once per step because the result is identical for every sample;
  * Gumbel uniforms come from the in-kernel counter-based generator (inject ``gumbel_uniform`` to
    reproduce a given draw); ``node_feats`` stays resident on the device instead of being re-uploaded.
"""
import pickle

import torch
from torch import nn

from step_b200 import ops

_NUM_NODES = {"METR-LA": 207, "PEMS04": 307, "PEMS03": 358, "PEMS-BAY": 325, "PEMS07": 883, "PEMS08": 170}
_TRAIN_LENGTH = {"METR-LA": 23990, "PEMS04": 13599, "PEMS03": 15303, "PEMS07": 16513, "PEMS-BAY": 36482, "PEMS08": 14284}
_DIM_FC = {"METR-LA": 383552, "PEMS04": 217296, "PEMS03": 244560, "PEMS07": 263920, "PEMS-BAY": 583424, "PEMS08": 228256}
_DIM_FC_MEAN = {"METR-LA": 16128, "PEMS-BAY": 16128, "PEMS03": 16128 * 2, "PEMS04": 16128 * 2, "PEMS07": 16128,
                "PEMS08": 16128 * 2}


def _load_pkl(path):
    with open(path, "rb") as f:
        try:
            return pickle.load(f)
        except UnicodeDecodeError:
            f.seek(0)
            return pickle.load(f, encoding="latin1")


class DiscreteGraphLearning(nn.Module):
    """Dynamic graph learning module."""

    def __init__(self, dataset_name, k, input_seq_len, output_seq_len):
        super().__init__()
        self.k = k
        self.num_nodes = _NUM_NODES[dataset_name]
        self.train_length = _TRAIN_LENGTH[dataset_name]
        data = _load_pkl("datasets/" + dataset_name + "/data_in{0}_out{1}.pkl".format(input_seq_len, output_seq_len))
        # plain attribute (not a buffer) so that the state dict matches the reference's
        self.node_feats = torch.from_numpy(data["processed_data"]).float()[:self.train_length, :, 0]
        self.dim_fc = _DIM_FC[dataset_name]
        self.embedding_dim = 100
        self.conv1 = nn.Conv1d(1, 8, 10, stride=1)
        self.conv2 = nn.Conv1d(8, 16, 10, stride=1)
        self.fc = nn.Linear(self.dim_fc, self.embedding_dim)
        self.bn1 = nn.BatchNorm1d(8)
        self.bn2 = nn.BatchNorm1d(16)
        self.bn3 = nn.BatchNorm1d(self.embedding_dim)
        self.dim_fc_mean = _DIM_FC_MEAN[dataset_name]
        self.fc_mean = nn.Linear(self.dim_fc_mean, 100)          # unused (as in the reference, :74,142)
        self.fc_cat = nn.Linear(self.embedding_dim, 2)
        self.fc_out = nn.Linear(self.embedding_dim * 2, self.embedding_dim)
        self.dropout = nn.Dropout(0.5)                            # unused (as in the reference)
        self.gumbel_uniform = None     # optional [B, N*N, 2] U(0,1) draws to inject (tests / reproduction)
        self.theta = None              # softmax(bernoulli_unnorm)[..., 0] of the last forward, [N, N]
        self.before_trainable = None   # optional callable run after the frozen encoder, before the first trainable parameter is read
        self._calls = 0
        self._feats_dev = None

    @staticmethod
    @torch.no_grad()
    def _update_running(bn, mean, var, count):
        bn.running_mean.mul_(1 - bn.momentum).add_(mean, alpha=bn.momentum)
        bn.running_var.mul_(1 - bn.momentum).add_(var, alpha=bn.momentum * count / max(count - 1, 1))
        bn.num_batches_tracked += 1

    @staticmethod
    def _eval_stats(bn):
        scale = bn.weight * torch.rsqrt(bn.running_var + bn.eps)
        return torch.stack([bn.running_mean, bn.running_var, scale, bn.bias - bn.running_mean * scale]).contiguous()

    def _global_feature(self, device):
        """Batch-invariant node embedding, reference :131-135.  [N, 100].  conv1/bn1/conv2/bn2 run in the fused
        trunk kernels (csrc/trunk.cu); the [N, dim_fc] x [dim_fc, 100] Linear + ReLU + bn3 in csrc/trunk_fc.cu."""
        if self._feats_dev is None or self._feats_dev.device != device:
            self._feats_dev = self.node_feats.to(device).t().contiguous()          # [N, L], uploaded once
        t = self.training
        e1 = None if t else self._eval_stats(self.bn1)
        e2 = None if t else self._eval_stats(self.bn2)
        y2n, s1, s2 = ops.TrunkConv.apply(self._feats_dev, self.conv1.weight, self.conv1.bias, self.bn1.weight, self.bn1.bias,
                                          self.conv2.weight, self.conv2.bias, self.bn2.weight, self.bn2.bias, self.bn1.eps, t, e1, e2)
        if t:
            n, L0 = self._feats_dev.shape
            self._update_running(self.bn1, s1[0], s1[1], n * (L0 - 9))
            self._update_running(self.bn2, s2[0], s2[1], n * (L0 - 18))
        # fc + ReLU + bn3: split-bf16 tcgen05 GEMMs (fp32-class accuracy in both precision modes), csrc/trunk_fc.cu
        e3 = None if t else torch.stack([self.bn3.running_mean, self.bn3.running_var]).contiguous()
        feat, s3 = ops.TrunkFc.apply(y2n, self.fc.weight, self.fc.bias, self.bn3.weight, self.bn3.bias, self.bn3.eps, t, e3, None)
        if t:
            self._update_running(self.bn3, s3[0], s3[1], y2n.shape[0])
        return feat

    def get_k_nn_neighbor(self, data, k=11 * 207, metric="cosine"):
        if metric != "cosine":
            raise NotImplementedError("only the cosine metric is used by STEP")
        with torch.no_grad():
            return ops.topk_mask(ops.cosine_gram(data), k)

    def forward(self, long_term_history, tsformer):
        """long_term_history [B, P*L, N, C] -> (bernoulli_unnorm [B,N*N,2], hidden [B,N,P,d], adj_knn, sampled_adj)."""
        batch_size, _, num_nodes, _ = long_term_history.shape
        # the frozen encoder first: it touches no trainable parameter, so a gradient all-reduce of the previous step that is
        # still in flight (parallel.GradReducer.reduce(async_op=True)) overlaps with it and is joined right after
        hidden_states = tsformer(long_term_history[..., 0:1])        # a strided view: the encoder reads it in place
        if self.before_trainable is not None:
            self.before_trainable()
        feat = self._global_feature(long_term_history.device)
        half = self.embedding_dim
        # the two halves of fc_out (reference :148-151 after the one-hot gathers), split-bf16 tcgen05 GEMMs
        w_send, w_recv = self.fc_out.weight[:, :half].contiguous(), self.fc_out.weight[:, half:].contiguous()
        ut = ops.Linear.apply(w_send, feat, None, False)                      # [100, N]  sender half (index j)
        v = ops.Linear.apply(feat, w_recv, self.fc_out.bias, False)           # [N, 100]  receiver half (index i)
        logits, theta = ops.EdgeLogits.apply(ut, v, self.fc_cat.weight, self.fc_cat.bias)
        self.theta = theta
        bernoulli_unnorm = logits.view(1, num_nodes * num_nodes, 2).expand(batch_size, -1, -1)
        self._calls += 1
        seed = (torch.initial_seed() + 0xC2B2AE35 * self._calls) & (2 ** 63 - 1)
        sampled_adj = ops.GumbelSample.apply(logits, self.gumbel_uniform, batch_size, 0.5, seed)
        seq_img = getattr(tsformer, "seq_image", None)
        if seq_img is not None:
            # bf16 mode: the encoder also emitted its output as the K-major operand image of the Gram GEMM (tcgen05)
            with torch.no_grad():
                shard = getattr(tsformer, "node_shard", None)
                if shard is not None and shard[1] > 1 and getattr(tsformer, "gathered_patches", None):
                    from step_b200 import parallel
                    sim = ops.tc_cosine_gram_sharded(seq_img, batch_size, num_nodes, tsformer.gathered_patches, shard[0], shard[1],
                                                     parallel.all_reduce_sum)
                else:
                    sim = ops.tc_cosine_gram(seq_img, batch_size, num_nodes, hidden_states.shape[2])
                adj_knn = ops.topk_mask(sim, self.k * self.num_nodes)
        else:
            adj_knn = ops.knn_prior(hidden_states, self.k * self.num_nodes)
        return bernoulli_unnorm, hidden_states, adj_knn, sampled_adj
