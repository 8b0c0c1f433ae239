"""Batch-parallel training plumbing: one process per GPU, NCCL over NVLink (gloo on CPU for tests).

Reference behaviour being reproduced: easytorch wraps the model in
DistributedDataParallel(find_unused_parameters=True) - every rank holds all parameters, runs its own batch
(batch size is per process), BatchNorm statistics stay per rank, and the gradients of the trainable parameters
are averaged across ranks once per step.  Parameters that receive no gradient (residual_convs.*, bn.7.*,
gconv.7.mlp.*, fc_mean.*) contribute zeros.

Implementation: all trainable gradients live in ONE flat buffer (each ``param.grad`` is a view into it), so the
collective is a single all-reduce of ~162 MB (METR-LA) with no gather/scatter copies; NVSwitch makes its cost
(~0.4 ms at the measured 725 GB/s bus bandwidth) small against the 30 ms step, so it is issued after backward
rather than bucketed."""
from __future__ import annotations

import os
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> tuple:
    """Initialise torch.distributed from RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun). Returns (rank, world, local_rank)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kwargs["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend, **kwargs)
    return rank, world, local_rank


def shard_batch(global_batch: int, rank: int, world: int) -> slice:
    """Contiguous slice of a global batch owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(global_batch, world)
    start = rank * base + min(rank, rem)
    return slice(start, start + base + (1 if rank < rem else 0))


class FlatGradReducer:
    """Keeps every trainable gradient in one flat buffer and averages it across ranks with a single all-reduce."""

    def __init__(self, params: Iterable[torch.nn.Parameter], world: Optional[int] = None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatGradReducer: no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, device=dev, dtype=dt)
        self.world = world if world is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self) -> None:
        """Call before backward: gradients accumulate into the views (parameters without a gradient stay zero)."""
        self.flat.zero_()

    def reduce(self) -> torch.Tensor:
        """Average the gradients over all ranks (no-op for a single process)."""
        if self.world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(self.world)
        return self.flat


class GradReducer:
    """Gradient averaging without accumulate kernels: ``zero()`` sets every ``param.grad`` to None, so autograd
    *assigns* the freshly computed gradient tensors (no ``grad += new`` launch per parameter, no memset of a 160 MB
    flat buffer per step).  ``reduce()``: tensors with at least ``big_numel`` elements (``discrete_graph_learning.fc.weight``
    is >= 95 % of STEP's gradient bytes) are all-reduced in place; all the small ones travel
    packed in one flat buffer (two multi-tensor copies).  Parameters that received no gradient stay None on every
    rank (the set is the same on all ranks: it is a property of the graph, not of the data)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], world: Optional[int] = None, big_numel: int = 1 << 20):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("GradReducer: no trainable parameters")
        self.world = world if world is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        self.big_numel = big_numel
        self._flat = None
        self._comm_stream = None
        self._pending = False

    def zero(self) -> None:
        for p in self.params:
            p.grad = None

    def reduce(self, async_op: bool = False) -> None:
        """Average the gradients over the ranks.  ``async_op=True``: the collectives are enqueued on a side stream and
        ``wait()`` joins them later - the frozen TSFormer encoder of the NEXT step (40 % of the step, no trainable
        parameter) runs meanwhile, so the 160 MB all-reduce leaves the critical path; ``wait()`` must be called before the
        gradients / the parameters they update are used (STEP: ``discrete_graph_learning.before_trainable``)."""
        if self.world <= 1:
            return
        big, small = [], []
        for p in self.params:
            g = p.grad
            if g is None:
                continue
            if not g.is_contiguous():
                g = g.contiguous()
                p.grad = g
            (big if g.numel() >= self.big_numel else small).append(g)
        dev = (big + small)[0].device if (big or small) else None
        if async_op and dev is not None and dev.type == "cuda":
            if self._comm_stream is None:
                self._comm_stream = torch.cuda.Stream(dev)
            self._comm_stream.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(self._comm_stream):
                self._reduce_now(big, small)
                for g in big + small:
                    g.record_stream(self._comm_stream)
            self._pending = True
        else:
            self._reduce_now(big, small)

    def _reduce_now(self, big, small) -> None:
        works = [dist.all_reduce(g, op=dist.ReduceOp.SUM, async_op=True) for g in big]
        if small:
            n = sum(g.numel() for g in small)
            if self._flat is None or self._flat.numel() != n or self._flat.device != small[0].device:
                self._flat = torch.empty(n, device=small[0].device, dtype=small[0].dtype)
            views, off = [], 0
            for g in small:
                views.append(self._flat[off:off + g.numel()].view_as(g))
                off += g.numel()
            torch._foreach_copy_(views, small)
            dist.all_reduce(self._flat, op=dist.ReduceOp.SUM)
            self._flat.div_(self.world)
            torch._foreach_copy_(small, views)
        for w in works:
            w.wait()
        if big:
            torch._foreach_div_(big, float(self.world))

    def wait(self) -> None:
        """Join an asynchronous ``reduce``: the current stream waits for the side stream's collectives."""
        if self._pending:
            torch.cuda.current_stream(self._comm_stream.device).wait_stream(self._comm_stream)
            self._pending = False


def node_shard_bounds(num_nodes: int, rank: int, world: int) -> tuple:
    """[n0, n1) of the nodes whose TSFormer sequences `rank` encodes in node-sharded mode."""
    sl = shard_batch(num_nodes, rank, world)
    return sl.start, sl.stop


def all_gather_nodes(local: torch.Tensor, num_nodes: int, rank: int, world: int, group=None) -> torch.Tensor:
    """Node-sharded mode (STEP_PEMS07): every rank encoded the sequences of its own node
    range; ONE all-gather over NVLink assembles the full hidden states [B, N, P, d] before the N x N similarity.
    `local`: [B, n1-n0, P, d].  Shards may differ by one node, so they travel padded to the largest shard."""
    if world == 1:
        return local
    B, nloc, P, d = local.shape
    nmax = (num_nodes + world - 1) // world
    send = local
    if nloc < nmax:
        send = torch.zeros(B, nmax, P, d, device=local.device, dtype=local.dtype)
        send[:, :nloc] = local
    gathered = torch.empty(world, B, nmax, P, d, device=local.device, dtype=local.dtype)
    dist.all_gather_into_tensor(gathered.view(world * B, nmax, P, d), send.contiguous(), group=group)
    full = torch.empty(B, num_nodes, P, d, device=local.device, dtype=local.dtype)
    for r in range(world):
        n0, n1 = node_shard_bounds(num_nodes, r, world)
        full[:, n0:n1] = gathered[r, :, : n1 - n0]
    return full


def gather_node_rows(local: torch.Tensor, num_nodes: int, rank: int, world: int, group=None) -> torch.Tensor:
    """All-gather along the node axis (dim 1) for any trailing shape: local [B, n1-n0, ...] -> [B, N, ...].
    Shards may differ by one node and travel padded to the largest."""
    if world == 1:
        return local
    B, nloc = local.shape[:2]
    tail = tuple(local.shape[2:])
    nmax = (num_nodes + world - 1) // world
    send = local.contiguous()
    if nloc < nmax:
        send = torch.zeros((B, nmax) + tail, device=local.device, dtype=local.dtype)
        send[:, :nloc] = local
    gathered = torch.empty((world, B, nmax) + tail, device=local.device, dtype=local.dtype)
    dist.all_gather_into_tensor(gathered.view((world * B, nmax) + tail), send, group=group)
    full = torch.empty((B, num_nodes) + tail, device=local.device, dtype=local.dtype)
    for r in range(world):
        n0, n1 = node_shard_bounds(num_nodes, r, world)
        full[:, n0:n1] = gathered[r, :, : n1 - n0]
    return full


def all_gather_seq_image(img_local: torch.Tensor, B: int, KC: int, num_nodes: int, rank: int, world: int, group=None) -> torch.Tensor:
    """Node-parallel mode, bf16 path: the encoder's Gram operand image of the LOCAL nodes ([B][KC][R_loc][8] bf16, as a flat
    uint8 tensor, R_loc = nodes rounded up to 128) is all-gathered over NVLink - 2 bytes per hidden value instead of the 4 of
    the fp32 states (114 vs 228 MB at PEMS07) - and re-assembled as the image of all N nodes ([B][KC][R_full][8])."""
    n0, n1 = node_shard_bounds(num_nodes, rank, world)
    nloc = n1 - n0
    r_loc = (nloc + 127) // 128 * 128
    r_full = (num_nodes + 127) // 128 * 128
    loc = img_local.view(B, KC, r_loc, 16)
    if world == 1:
        return img_local
    nmax = (num_nodes + world - 1) // world
    send = loc[:, :, :nmax] if r_loc >= nmax else torch.cat(
        [loc, torch.zeros(B, KC, nmax - r_loc, 16, device=loc.device, dtype=loc.dtype)], dim=2)
    send = send.contiguous()
    gathered = torch.empty(world, B, KC, nmax, 16, device=loc.device, dtype=loc.dtype)
    dist.all_gather_into_tensor(gathered.view(world * B, KC, nmax, 16), send, group=group)
    full = torch.zeros(B, KC, r_full, 16, device=loc.device, dtype=loc.dtype)
    for r in range(world):
        m0, m1 = node_shard_bounds(num_nodes, r, world)
        full[:, :, m0:m1] = gathered[r, :, :, : m1 - m0]
    return full.view(-1)


def all_reduce_sum(t: torch.Tensor, group=None) -> torch.Tensor:
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t
