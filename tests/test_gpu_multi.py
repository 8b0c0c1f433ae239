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
    model.node_shard = (rank, world)
    got = model(x)
    ok = bool(torch.equal(ref, got))
    err = float((ref - got).abs().max())
    img_ok = True
    if precision == "bf16":
        from step_b200 import ops
        sim_a = ops.tc_cosine_gram(model.seq_image, B, N, P)
        model.node_shard = None
        model(x)
        sim_b = ops.tc_cosine_gram(model.seq_image, B, N, P)
        img_ok = bool(torch.equal(sim_a, sim_b))
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
