"""CPU, world_size 2 over gloo: the N>1 host logic (flat-gradient all-reduce, batch sharding)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from step_b200 import parallel
    r, w, _ = parallel.init_from_env("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    unused = torch.nn.Parameter(torch.ones(5))          # never receives a gradient (like fc_mean.* in STEP)
    frozen = torch.nn.Parameter(torch.ones(2), requires_grad=False)
    red = parallel.FlatGradReducer(list(model.parameters()) + [unused, frozen])
    assert red.numel == sum(p.numel() for p in model.parameters()) + 5
    x = torch.arange(8, dtype=torch.float32).view(2, 4) + 10.0 * rank      # different data per rank
    for _ in range(2):                                                       # second pass checks zero()
        red.zero()
        model(x).square().sum().backward()
        flat = red.reduce().clone()
    # reference: average of the two ranks' gradients computed locally
    grads = []
    for rr in range(world):
        m2 = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
        m2.load_state_dict(model.state_dict())
        xx = torch.arange(8, dtype=torch.float32).view(2, 4) + 10.0 * rr
        m2(xx).square().sum().backward()
        grads.append(torch.cat([p.grad.reshape(-1) for p in m2.parameters()]))
    want = torch.cat([(grads[0] + grads[1]) / 2, torch.zeros(5)])
    ok = torch.allclose(flat, want, rtol=1e-5, atol=1e-5) and unused.grad.abs().sum().item() == 0.0
    # views: param.grad aliases the flat buffer
    ok = ok and model[0].weight.grad.data_ptr() == red.flat.data_ptr()
    out.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_grad_reducer_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


def _worker_assign(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from step_b200 import parallel
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    unused = torch.nn.Parameter(torch.ones(5))
    frozen = torch.nn.Parameter(torch.ones(2), requires_grad=False)
    red = parallel.GradReducer(list(model.parameters()) + [unused, frozen], big_numel=10)   # Linear(4,3).weight is "big"
    x = torch.arange(8, dtype=torch.float32).view(2, 4) + 10.0 * rank
    for _ in range(2):
        red.zero()
        assert all(p.grad is None for p in model.parameters())
        model(x).square().sum().backward()
        red.reduce()
    got = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    grads = []
    for rr in range(world):
        m2 = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
        m2.load_state_dict(model.state_dict())
        xx = torch.arange(8, dtype=torch.float32).view(2, 4) + 10.0 * rr
        m2(xx).square().sum().backward()
        grads.append(torch.cat([p.grad.reshape(-1) for p in m2.parameters()]))
    ok = torch.allclose(got, (grads[0] + grads[1]) / 2, rtol=1e-5, atol=1e-5) and unused.grad is None
    out.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_grad_reducer_assign_mode_world2():
    """GradReducer: gradients assigned by autograd (grad = None before backward), big tensors all-reduced in place,
    small ones packed; result = the average over ranks; parameters without a gradient stay None."""
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_assign, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


def test_shard_batch():
    from step_b200.parallel import shard_batch
    for gb, w in ((32, 8), (33, 8), (4, 8), (7, 2)):
        parts = [shard_batch(gb, r, w) for r in range(w)]
        assert parts[0].start == 0 and parts[-1].stop == gb
        assert all(a.stop == b.start for a, b in zip(parts, parts[1:]))
        sizes = [p.stop - p.start for p in parts]
        assert max(sizes) - min(sizes) <= 1


def _gather_worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from step_b200 import parallel
    parallel.init_from_env("gloo")
    B, N, P, d = 2, 7, 3, 4                        # 7 nodes over 2 ranks: shards of 4 and 3
    full = torch.arange(B * N * P * d, dtype=torch.float32).view(B, N, P, d)
    n0, n1 = parallel.node_shard_bounds(N, rank, world)
    got = parallel.all_gather_nodes(full[:, n0:n1].contiguous(), N, rank, world)
    out.put((rank, bool(torch.equal(got, full)), (n0, n1)))
    dist.barrier()
    dist.destroy_process_group()


def test_all_gather_nodes_world2_uneven():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True, (0, 4)), (1, True, (4, 7))]


def _image_worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from step_b200 import parallel
    parallel.init_from_env("gloo")
    B, N, KC = 2, 300, 5                            # 300 nodes over 2 ranks: 150 + 150 -> local images padded to 256 rows
    r_full = (N + 127) // 128 * 128
    g = torch.Generator().manual_seed(3)
    full = torch.randint(0, 255, (B, KC, r_full, 16), dtype=torch.uint8, generator=g)
    full[:, :, N:] = 0
    n0, n1 = parallel.node_shard_bounds(N, rank, world)
    r_loc = (n1 - n0 + 127) // 128 * 128
    local = torch.zeros(B, KC, r_loc, 16, dtype=torch.uint8)
    local[:, :, : n1 - n0] = full[:, :, n0:n1]
    got = parallel.all_gather_seq_image(local.view(-1), B, KC, N, rank, world)
    ok_img = bool(torch.equal(got.view(B, KC, r_full, 16), full))
    # node rows with an arbitrary trailing shape (the last-patch hidden states [B, n, 1, 96]) and an uneven split
    N2 = 7
    rows = torch.arange(B * N2 * 1 * 4, dtype=torch.float32).view(B, N2, 1, 4)
    m0, m1 = parallel.node_shard_bounds(N2, rank, world)
    ok_rows = bool(torch.equal(parallel.gather_node_rows(rows[:, m0:m1], N2, rank, world), rows))
    # the Gram exchange: every rank fills its tiles' rows of a zero matrix, the sum over ranks is the full matrix
    gram = torch.zeros(3, 3)
    gram[rank] = float(rank + 1)
    parallel.all_reduce_sum(gram)
    ok_sum = bool(torch.equal(gram, torch.tensor([[1.0] * 3, [2.0] * 3, [0.0] * 3])))
    out.put((rank, ok_img, ok_rows, ok_sum))
    dist.barrier()
    dist.destroy_process_group()


def test_node_parallel_exchanges_world2():
    """Node-parallel mode host logic: bf16 sequence-image all-gather + re-assembly, last-patch row gather, Gram row exchange."""
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_image_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True, True, True), (1, True, True, True)]
