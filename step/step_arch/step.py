"""STEP = frozen TSFormer + discrete graph learning + Graph WaveNet, on the B200-native kernels.

Drop-in for ``step/step_arch/step.py:9-72`` of the reference: same constructor, ``forward`` signature and
4-tuple return, same state-dict keys (``tsformer.*``, ``backend.*``, ``discrete_graph_learning.*``).
"""
import torch
from torch import nn

from .tsformer import TSFormer
from .graphwavenet import GraphWaveNet
from .discrete_graph_learning import DiscreteGraphLearning


class STEP(nn.Module):
    def __init__(self, dataset_name, pre_trained_tsformer_path, tsformer_args, backend_args, dgl_args):
        super().__init__()
        self.dataset_name = dataset_name
        self.pre_trained_tsformer_path = pre_trained_tsformer_path
        self.tsformer = TSFormer(**tsformer_args)
        self.backend = GraphWaveNet(**backend_args)
        self.load_pre_trained_model()
        self.discrete_graph_learning = DiscreteGraphLearning(**dgl_args)

    def load_pre_trained_model(self):
        """Load and freeze the pre-trained TSFormer (reference step.py:27-35; checkpoints were saved from
        CUDA, so map them to CPU first - parameters move with the module afterwards)."""
        checkpoint_dict = torch.load(self.pre_trained_tsformer_path, map_location="cpu")
        self.tsformer.load_state_dict(checkpoint_dict["model_state_dict"])
        for param in self.tsformer.parameters():
            param.requires_grad = False

    def forward(self, history_data: torch.Tensor, long_history_data: torch.Tensor, future_data: torch.Tensor,
                batch_seen: int, epoch: int, **kwargs):
        """history_data [B,12,N,3], long_history_data [B,P*12,N,3] ->
        (y_hat [B,12,N,1], theta [B,N,N], adj_knn [B,N,N], gsl_coefficient)."""
        batch_size, _, num_nodes, _ = history_data.shape
        bernoulli_unnorm, hidden_states, adj_knn, sampled_adj = self.discrete_graph_learning(long_history_data, self.tsformer)
        hidden_last = hidden_states[:, :, -1, :]
        y_hat = self.backend(history_data, hidden_states=hidden_last, sampled_adj=sampled_adj).transpose(1, 2)
        gsl_coefficient = 1 / (int(epoch / 6) + 1) if epoch is not None else 0
        # softmax(bernoulli_unnorm)[..., 0] is identical for every sample (the logits are batch-invariant)
        theta = self.discrete_graph_learning.theta.unsqueeze(0).expand(batch_size, num_nodes, num_nodes)
        return y_hat.unsqueeze(-1), theta, adj_knn, gsl_coefficient
