"""STEP = frozen TSFormer + discrete graph learning + Graph WaveNet, on the B200-native kernels.

Drop-in for ``step/step_arch/step.py:9-72`` of the reference: same constructor, ``forward`` signature and
4-tuple return, same state-dict keys (``tsformer.*``, ``backend.*``, ``discrete_graph_learning.*``).
"""
from typing import Optional, Tuple

import torch
from torch import nn

from .discrete_graph_learning import DiscreteGraphLearning
from .graphwavenet import GraphWaveNet
from .tsformer import TSFormer

_GSL_DECAY_EPOCHS = 6      # the graph-structure-learning loss weight is 1 / (epoch // 6 + 1)  (reference step.py:64-67)


def gsl_weight(epoch: Optional[int]) -> float:
    """Weight of the BCE(theta, kNN prior) regulariser; 0 when the caller passes no epoch (evaluation)."""
    return 0 if epoch is None else 1 / (int(epoch / _GSL_DECAY_EPOCHS) + 1)


class STEP(nn.Module):
    """Sub-module attribute names are part of the checkpoint contract: ``tsformer``, ``backend``,
    ``discrete_graph_learning``."""

    def __init__(self, dataset_name, pre_trained_tsformer_path, tsformer_args, backend_args, dgl_args):
        super().__init__()
        self.dataset_name, self.pre_trained_tsformer_path = dataset_name, pre_trained_tsformer_path
        self.tsformer = TSFormer(**tsformer_args)
        self.backend = GraphWaveNet(**backend_args)
        self.load_pre_trained_model()
        self.discrete_graph_learning = DiscreteGraphLearning(**dgl_args)

    def load_pre_trained_model(self) -> None:
        """Stage-1 weights in, gradients off (reference step.py:27-35).  The shipped checkpoints were pickled from
        CUDA tensors, hence ``map_location``; the parameters follow the module to its device afterwards."""
        state = torch.load(self.pre_trained_tsformer_path, map_location="cpu")["model_state_dict"]
        self.tsformer.load_state_dict(state)
        self.tsformer.requires_grad_(False)

    def forward(self, history_data: torch.Tensor, long_history_data: torch.Tensor, future_data: torch.Tensor,
                batch_seen: int, epoch: int, **kwargs) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, float]:
        """history_data [B,12,N,3], long_history_data [B,P*12,N,3] ->
        (y_hat [B,12,N,1], theta [B,N,N], adj_knn [B,N,N], gsl_coefficient)."""
        B, N = history_data.shape[0], history_data.shape[2]
        graph = self.discrete_graph_learning
        _, hidden, prior_graph, sampled_graph = graph(long_history_data, self.tsformer)
        # only the representation of the most recent patch conditions the forecaster (reference step.py:58)
        forecast = self.backend(history_data, hidden_states=hidden[:, :, -1, :], sampled_adj=sampled_graph)   # [B, N, 12]
        # theta = softmax(edge logits)[..., 0] is batch-invariant: the module keeps it as [N, N]; expose the
        # reference's [B, N, N] shape as a stride-0 view (the fused loss recognises it)
        theta = graph.theta.unsqueeze(0).expand(B, N, N)
        return forecast.transpose(1, 2).unsqueeze(-1), theta, prior_graph, gsl_weight(epoch)
