"""Multi-GPU (>= 2 devices) tests: node-sharded TSFormer encoding with one NCCL all-gather must reproduce the
single-GPU hidden states.  Skipped on single-GPU boxes; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, precision, out):
    import torch.distributed as dist
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    os.environ["NCCL_DEBUG"] = "WARN"
    import sys
    sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
    from conftest import TS_ARGS
    from oracle import step_oracle as O
    from step.step_arch import TSFormer
    from step_b200 import parallel
    parallel.init_from_env("nccl")
    dev = torch.device("cuda", rank)
    model = TSFormer(**TS_ARGS)
    model.load_state_dict(O.synthetic_tsformer_params(3), strict=True)
    model = model.to(dev).eval()
    model.precision = precision
    g = torch.Generator().manual_seed(11)
    B, N, P = 2, 45, 168                                  # 45 nodes over 2 ranks: 23 + 22
    x = torch.randn(B, P * 12, N, 1, generator=g).to(dev)
    ref = model(x)
    ref_img = None if model.seq_image is None else model.seq_image.clone()
    model.node_shard = (rank, world)
    got = model(x)
    img_ok = True
    if precision == "bf16":
        # bf16 path: the bf16 Gram operand image travels (half the bytes), hidden carries the last patch only
        from step_b200 import ops
        want = ref[:, :, -1:, :]
        ok = bool(torch.equal(want, got)) and bool(torch.equal(model.seq_image, ref_img))
        err = float((want - got).abs().max())
        sim_a = ops.tc_cosine_gram_sharded(model.seq_image, B, N, P, rank, world, parallel.all_reduce_sum)
        sim_b = ops.tc_cosine_gram(ref_img, B, N, P)
        img_ok = bool(torch.equal(sim_a, sim_b))
    else:
        ok = bool(torch.equal(ref, got))
        err = float((ref - got).abs().max())
    out.put((rank, ok, err, img_ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_node_sharded_encoder_matches_single_gpu(precision):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, precision, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, ok, err, img_ok in res:
        assert ok, (rank, err)
        assert img_ok, rank


def _model_worker(rank, world, port, tmp, out):
    import torch.distributed as dist
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    os.environ["NCCL_DEBUG"] = "WARN"
    import sys
    import pathlib
    sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from conftest import build_step_model
    from oracle import step_oracle as O
    from step.step_loss import step_loss
    from step_b200 import parallel
    parallel.init_from_env("nccl")
    dev = torch.device("cuda", rank)
    ds = "PEMS08"
    model, _, _ = build_step_model(pathlib.Path(tmp) / f"r{rank}", ds, 0, real_ckpt=False)
    model = model.to(dev).train()
    model.tsformer.dropout_p = 0.0
    model.backend.dropout = 0.0
    history, long_history, future, uniform = O.synthetic_batch(ds, 2, 168, 4)
    history, long_history, future = history.to(dev), long_history.to(dev), future.to(dev)
    model.discrete_graph_learning.gumbel_uniform = uniform.to(dev)
    res = []
    for shard in (None, (rank, world)):
        model.tsformer.node_shard = shard
        for p in model.parameters():
            p.grad = None
        y, theta, knn, coeff = model(history_data=history, long_history_data=long_history, future_data=None, batch_seen=0, epoch=1)
        loss = step_loss(y[..., :1], future[..., :1], theta, knn, coeff, null_val=0.0)
        loss.backward()
        res.append((y.detach().clone(), knn.clone(), model.backend.fc_his[0].weight.grad.clone(),
                    model.discrete_graph_learning.fc.weight.grad.clone()))
    ok = all(bool(torch.equal(a, b)) for a, b in zip(res[0][:2], res[1][:2]))
    gerr = max(float((a - b).abs().max() / a.abs().max().clamp_min(1e-12)) for a, b in zip(res[0][2:], res[1][2:]))
    out.put((rank, ok, gerr))
    dist.barrier()
    dist.destroy_process_group()


def test_node_parallel_step_matches_single_gpu(tmp_path):
    """BASELINE configs[4] mode on 2 GPUs: node-sharded TSFormer + bf16 image all-gather + sharded Gram rows; the whole STEP
    forward (y_hat, kNN graph) is bit-identical to the single-GPU pass, gradients equal up to atomic summation order."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_model_worker, args=(r, 2, port, str(tmp_path), out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, ok, gerr in res:
        assert ok, rank
        assert gerr < 1e-4, (rank, gerr)
