"""Graph WaveNet backbone on the B200-native kernels (drop-in for the reference module).

Constructor / ``forward`` signature and state-dict keys follow
``step/step_arch/graphwavenet/model.py:51-224`` of the reference.  The eight gated-TCN + diffusion-GCN
layers run in ``step_b200.ops.GWNetStack`` (hand-written forward and backward); the prologue (2->32 start
conv, support normalisation, adaptive adjacency: ``csrc/gw_glue.cu``) and the epilogue (fc_his, end convs:
split-bf16 tcgen05 GEMMs, ``csrc/tc_gemm.cu``) are autograd Functions over the same C ABI - no library GEMM
or eager elementwise kernel is left on the path.
"""
import torch
from torch import nn

from step_b200 import ops


class _Conv1x1(nn.Module):
    """Parameter holder named like the reference's ``linear`` (model.py:18-24): ``.mlp`` is a 1x1 Conv2d."""

    def __init__(self, c_in, c_out):
        super().__init__()
        self.mlp = nn.Conv2d(c_in, c_out, kernel_size=(1, 1), bias=True)


class _GcnParams(nn.Module):
    """Parameter holder named like the reference's ``gcn`` (model.py:26-48): ``.mlp.mlp.{weight,bias}``."""

    def __init__(self, c_in, c_out, dropout, support_len=3, order=2):
        super().__init__()
        self.mlp = _Conv1x1((order * support_len + 1) * c_in, c_out)
        self.dropout = dropout
        self.order = order


class GraphWaveNet(nn.Module):
    def __init__(self, num_nodes, support_len, dropout=0.3, gcn_bool=True, addaptadj=True, aptinit=None, in_dim=2,
                 out_dim=12, residual_channels=32, dilation_channels=32, skip_channels=256, end_channels=512,
                 kernel_size=2, blocks=4, layers=2, **kwargs):
        super().__init__()
        if not (gcn_bool and addaptadj and aptinit is None and in_dim == 2 and residual_channels == 32
                and dilation_channels == 32 and skip_channels == 256 and kernel_size == 2 and layers == 2
                and blocks * layers <= 8 and support_len == 2):
            raise NotImplementedError("step_b200 GraphWaveNet kernels are specialised for the STEP backend_args "
                                      "(gcn + adaptive adjacency, 32/32/256 channels, kernel 2, 4 blocks x 2 layers)")
        self.dropout, self.blocks, self.layers, self.gcn_bool, self.addaptadj = dropout, blocks, layers, gcn_bool, addaptadj
        self.filter_convs, self.gate_convs = nn.ModuleList(), nn.ModuleList()
        self.residual_convs, self.skip_convs = nn.ModuleList(), nn.ModuleList()
        self.bn, self.gconv = nn.ModuleList(), nn.ModuleList()
        self.fc_his = nn.Sequential(nn.Linear(96, 512), nn.ReLU(), nn.Linear(512, 256), nn.ReLU())
        self.start_conv = nn.Conv2d(in_dim, residual_channels, kernel_size=(1, 1))
        self.supports_len = support_len + 1
        self.nodevec1 = nn.Parameter(torch.randn(num_nodes, 10), requires_grad=True)
        self.nodevec2 = nn.Parameter(torch.randn(10, num_nodes), requires_grad=True)
        receptive_field = 1
        for _ in range(blocks):
            additional_scope, new_dilation = kernel_size - 1, 1
            for _ in range(layers):
                self.filter_convs.append(nn.Conv2d(residual_channels, dilation_channels, (1, kernel_size), dilation=new_dilation))
                self.gate_convs.append(nn.Conv2d(residual_channels, dilation_channels, (1, kernel_size), dilation=new_dilation))
                self.residual_convs.append(nn.Conv2d(dilation_channels, residual_channels, (1, 1)))   # unused (as in the reference)
                self.skip_convs.append(nn.Conv2d(dilation_channels, skip_channels, (1, 1)))
                self.bn.append(nn.BatchNorm2d(residual_channels))
                new_dilation *= 2
                receptive_field += additional_scope
                additional_scope *= 2
                self.gconv.append(_GcnParams(dilation_channels, residual_channels, dropout, support_len=self.supports_len))
        self.end_conv_1 = nn.Conv2d(skip_channels, end_channels, (1, 1), bias=True)
        self.end_conv_2 = nn.Conv2d(end_channels, out_dim, (1, 1), bias=True)
        self.receptive_field = receptive_field
        self._calls = 0

    @staticmethod
    def _random_walk(adj):
        """D^-1 (A + I), reference model.py:121-130 (row sums >= 1, so no division guard is needed)."""
        a = adj + torch.eye(adj.shape[1], device=adj.device, dtype=adj.dtype)
        return a / a.sum(2, keepdim=True)

    def _flat_layer_params(self):
        n = self.blocks * self.layers
        flat = []
        for i in range(n):
            dead = (i == n - 1)      # last layer: gcn + bn outputs are discarded (reference model.py:217-218)
            flat += [self.filter_convs[i].weight, self.filter_convs[i].bias, self.gate_convs[i].weight, self.gate_convs[i].bias,
                     self.skip_convs[i].weight.view(256, 32), self.skip_convs[i].bias,
                     None if dead else self.gconv[i].mlp.mlp.weight.view(32, 224), None if dead else self.gconv[i].mlp.mlp.bias,
                     None if dead else self.bn[i].weight, None if dead else self.bn[i].bias]
        return flat

    def _eval_bn_stats(self, n, device):
        st = torch.zeros(n, 4, 32, device=device)
        for i in range(n - 1):
            bn = self.bn[i]
            scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
            st[i, 0], st[i, 1], st[i, 2], st[i, 3] = bn.running_mean, bn.running_var, scale, bn.bias - bn.running_mean * scale
        return st

    @torch.no_grad()
    def _update_running_stats(self, bn_stats, batch, num_nodes):
        """BatchNorm2d running statistics of the 7 live layers from the fused stack's batch statistics
        (momentum update with the unbiased variance, as nn.BatchNorm2d does) - multi-tensor ops: 6 launches."""
        n = self.blocks * self.layers - 1
        key = (batch, num_nodes, bn_stats.device)
        if getattr(self, "_rs_key", None) != key:          # unbiased-variance factors M / (M - 1), cached on the device
            t, factors = 13, []
            for i in range(n):
                t -= 1 if i % 2 == 0 else 2
                m = float(batch * t * num_nodes)
                factors.append(m / max(m - 1.0, 1.0))
            self._rs_factors = torch.tensor(factors, device=bn_stats.device, dtype=bn_stats.dtype).view(n, 1)
            self._rs_key = key
        mom = self.bn[0].momentum
        var_unb = bn_stats[:n, 1] * self._rs_factors
        means, variances = list(bn_stats[:n, 0].unbind(0)), list(var_unb.unbind(0))
        rms, rvs = [self.bn[i].running_mean for i in range(n)], [self.bn[i].running_var for i in range(n)]
        torch._foreach_mul_(rms, 1 - mom)
        torch._foreach_add_(rms, means, alpha=mom)
        torch._foreach_mul_(rvs, 1 - mom)
        torch._foreach_add_(rvs, variances, alpha=mom)
        torch._foreach_add_([self.bn[i].num_batches_tracked for i in range(n)], 1)

    def forward(self, input, hidden_states, sampled_adj):
        """input [B, L, N, C], hidden_states [B, N, 96], sampled_adj [B, N, N] -> [B, N, 12]."""
        B, L, N, _ = input.shape
        if L + 1 != self.receptive_field:
            raise NotImplementedError("step_b200 GraphWaveNet expects 12 history steps (receptive field 13)")
        n = self.blocks * self.layers
        # prologue (csrc/gw_glue.cu): start conv on the left-padded history straight into [B,13,N,32], both random-walk
        # supports of the sampled graph, the adaptive adjacency
        x0 = ops.GwStart.apply(input, self.start_conv.weight, self.start_conv.bias)
        P1, P2 = ops.GwSupports.apply(sampled_adj)
        P3 = ops.GwAdaptive.apply(self.nodevec1, self.nodevec2)
        self._calls += 1
        seed = (torch.initial_seed() + 0x85EBCA77 * self._calls) & (2 ** 63 - 1)
        training = self.training
        eval_stats = None if training else self._eval_bn_stats(n, input.device)
        skip, bn_stats = ops.GWNetStack.apply(x0, P1, P2, P3, training, self.dropout if training else 0.0, seed,
                                              eval_stats, n, *self._flat_layer_params())
        if training:
            self._update_running_stats(bn_stats, B, N)
        # epilogue (csrc/tc_gemm.cu): fc_his, skip add, end convs as split-bf16 tcgen05 GEMMs with fused bias/ReLU epilogues
        out = ops.GwEpilogue.apply(hidden_states.reshape(B * N, -1), skip.view(B * N, 256),
                                   self.fc_his[0].weight, self.fc_his[0].bias, self.fc_his[2].weight, self.fc_his[2].bias,
                                   self.end_conv_1.weight.view(512, 256), self.end_conv_1.bias,
                                   self.end_conv_2.weight.view(-1, 512), self.end_conv_2.bias)
        return out.view(B, N, -1)                                         # [B,N,12]
