#!/usr/bin/env python
"""STEP fwd+bwd throughput benchmark (BASELINE.json metric) - see the contract in DESIGN.md section 6.

  python bench.py --gpus 1 --steps 10 --warmup 3            # our B200-native path
  python bench.py --impl reference --steps 2 --warmup 1     # the reference algorithm on the host CPU cores
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         bench.py --gpus N --steps K --warmup W               # one rank per GPU, batch-parallel

A "step" = one training step of STEP_METR-LA (N=207 nodes, per-GPU batch 32, 168 patches of 12 = 2016-step
long history): forward (frozen TSFormer in train() exactly as the reference runs it, discrete graph
learning, Graph WaveNet), step_loss, backward to every trainable parameter (+ NCCL gradient all-reduce
when N > 1).  Synthetic N(0,1) inputs, real pre-trained TSFormer weights (tests/golden fixture), seeded
random GWNet/DGL weights.  Prints ONE JSON line.
"""
import argparse
import json
import os
import pickle
import subprocess
import sys
import tempfile
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DATASET = "METR-LA"
NODES, BATCH, PATCHES = 207, 32, 168
METRIC = "STEP fwd+bwd samples/sec (STEP_METR-LA, N=207, per-GPU batch 32, 12->12)"
# (nodes, per-GPU batch, patches) of the reference's STEP_<NAME>.py configs
WORKLOADS = {"METR-LA": (207, 32, 168), "PEMS04": (307, 8, 336), "PEMS-BAY": (325, 32, 168), "PEMS07": (883, 4, 168)}


def set_workload(name):
    global DATASET, NODES, BATCH, PATCHES, METRIC
    DATASET = name
    NODES, BATCH, PATCHES = WORKLOADS[name]
    METRIC = "STEP fwd+bwd samples/sec (STEP_%s, N=%d, per-GPU batch %d, 12->12)" % (name, NODES, BATCH)
    GW_ARGS["num_nodes"] = NODES
    TS_ARGS["num_token"] = float(PATCHES)


TS_ARGS = dict(patch_size=12, in_channel=1, embed_dim=96, num_heads=4, mlp_ratio=4, dropout=0.1, num_token=168.0,
               mask_ratio=0.75, encoder_depth=4, decoder_depth=1, mode="forecasting")
GW_ARGS = dict(num_nodes=NODES, support_len=2, dropout=0.3, gcn_bool=True, addaptadj=True, aptinit=None, in_dim=2,
               out_dim=12, residual_channels=32, dilation_channels=32, skip_channels=256, end_channels=512,
               kernel_size=2, blocks=4, layers=2)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default: the config's, 32 for METR-LA)")
    ap.add_argument("--workload", default="METR-LA", choices=sorted(WORKLOADS),
                    help="STEP config to run; METR-LA is the headline (BASELINE.json configs[1]), the others are "
                         "the remaining configs' shapes with synthetic TSFormer weights (not bench lines)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--only-resident", action="store_true", help="profiling aid: run only the device-resident loop")
    ap.add_argument("--no-dropout", action="store_true", help="parity-style run (all dropout off); not the headline")
    ap.add_argument("--precision", default=os.environ.get("STEP_B200_PRECISION", "bf16"), choices=["bf16", "fp32"],
                    help="TSFormer encoder kernels: bf16 tcgen05 tensor cores (default, BASELINE config) or fp32 CUDA cores")
    ap.add_argument("--chunk-seqs", type=int, default=int(os.environ.get("STEP_B200_TS_CHUNK", "0")))
    return ap.parse_args()


# --------------------------------------------------------------------------------------------- helpers
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        time.sleep(0.05)
        sm = sorted(int(r[1]) for r in self.rows if len(r) >= 8 and r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if len(r) >= 8 and r[2].isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def write_dataset(tmp, node_feats):
    d = os.path.join(tmp, "datasets", DATASET)
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "data_in12_out12.pkl"), "wb") as f:
        pickle.dump({"processed_data": node_feats.unsqueeze(-1).numpy()}, f)


def ts_state():
    if DATASET != "METR-LA":        # only the METR-LA checkpoint is small enough to ship as a fixture
        from step_b200 import synth
        return synth.synthetic_tsformer_params(1)
    return torch.load(os.path.join(ROOT, "tests", "golden", "tsformer_METR-LA_state.pt"))


# --------------------------------------------------------------------------------------------- CPU comparator
def cpu_reference_run(steps, warmup, batch, dropout=True):
    """The reference algorithm (oracle/step_oracle.py restatement of the reference's torch modules) on the
    host cores, all threads, same config except a bounded per-step batch.  Returns (samples/s, info)."""
    from oracle import step_oracle as O
    # torch CPU ops on these shapes slow down badly past ~16-32 threads (measured on the 128-thread GPU host:
    # 83 s/step with 128 threads); use the best-performing setting and report it as `cores`
    torch.set_num_threads(min(os.cpu_count() or 1, int(os.environ.get("STEP_B200_CPU_THREADS", "32"))))
    params = O.synthetic_trainable_params(DATASET, 0)
    sd = dict(params)
    sd.update(O.bn_buffers(DATASET))
    sd.update({"tsformer." + k: v for k, v in ts_state().items()})
    for k in params:
        sd[k] = sd[k].clone().requires_grad_(True)
    node_feats = O.synthetic_node_feats(DATASET, 0)
    history, long_history, future, uniform = O.synthetic_batch(DATASET, batch, PATCHES, 0)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        loss, _ = O.train_step(sd, history, long_history, future, node_feats, uniform, epoch=1, null_val=0.0,
                               gw_drop=0.3 if dropout else 0.0, ts_drop=0.1 if dropout else 0.0)
        loss.backward()
        for k in params:
            sd[k].grad = None
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    return batch / dt, {"cores": torch.get_num_threads(), "kind": "port",
                        "sample": f"{steps} timed fwd+bwd steps (after {warmup} warm-up) of STEP_{DATASET} at batch {batch} "
                                  f"(CPU samples/s is ~flat in batch), fp32, dropout {'live as in the reference train()' if dropout else 'off'}, "
                                  f"{dt:.2f} s/step"}


# --------------------------------------------------------------------------------------------- our arm
def main():
    args = parse()
    set_workload(args.workload)
    if args.batch is None:
        args.batch = BATCH
    if not os.environ.get("STEP_B200_KEEP_NCCL_DEBUG"):
        os.environ["NCCL_DEBUG"] = "WARN"       # NCCL's version banner goes to stdout and would precede the JSON line
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        b = 2 if (args.steps + args.warmup) <= 12 else 1
        v, info = cpu_reference_run(args.steps, args.warmup, b)
        info["value"] = v
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": "samples/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * b / v,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32",
                          "data": "synthetic", "config": {"workload": "STEP_%s N=%d P=%d 12->12, CPU sample batch %d" % (DATASET, NODES, PATCHES, b)},
                          "cpu_baseline": info,
                          "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the B200-native path has no CPU fallback "
                         "(use --impl reference for the CPU comparator)")
    import torch.distributed as dist
    from step_b200 import parallel
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    parallel.init_from_env("nccl")

    from step_b200 import synth as O             # deterministic synthetic tensors (the oracle is imported by the CPU leg only)
    from step.step_arch import STEP
    from step.step_loss import step_loss
    from step_b200 import ops

    B = args.batch
    node_feats = O.synthetic_node_feats(DATASET, 0)
    tmp = tempfile.mkdtemp(prefix="step_bench_")
    write_dataset(tmp, node_feats)
    torch.save({"model_state_dict": ts_state()}, os.path.join(tmp, "ts.pt"))
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        model = STEP(DATASET, os.path.join(tmp, "ts.pt"), dict(TS_ARGS), dict(GW_ARGS),
                     dict(dataset_name=DATASET, k=10, input_seq_len=12, output_seq_len=12))
    finally:
        os.chdir(cwd)
    full = dict(O.synthetic_trainable_params(DATASET, 0))
    model.load_state_dict(full, strict=False)
    model = model.to(dev).train()                 # the reference trains with the frozen TSFormer left in train()
    model.tsformer.chunk_seqs = args.chunk_seqs
    model.tsformer.precision = args.precision
    if args.no_dropout:
        model.tsformer.dropout_p = 0.0
        model.backend.dropout = 0.0
    # gradients are assigned (not accumulated) by autograd; fc.weight is all-reduced in place, the rest packed
    if os.environ.get("STEP_B200_REDUCER", "assign") == "flat":
        reducer = parallel.FlatGradReducer(model.parameters(), world)     # single flat buffer, gradients accumulate into views
    else:
        reducer = parallel.GradReducer(model.parameters(), world)

    torch.manual_seed(1234 + rank)
    n_host = 4                                    # rotate a few distinct host batches (inputs differ step to step)
    host = []
    for i in range(n_host):
        h, lh, f, _ = O.synthetic_batch(DATASET, B, PATCHES, 100 + 17 * rank + i)
        host.append((h.pin_memory(), lh.pin_memory(), f.pin_memory()))
    resident = [(h.to(dev), lh.to(dev), f.to(dev)) for (h, lh, f) in host[:2]]
    h2d_bytes = sum(t.numel() * 4 for t in host[0])

    def train_step(history, long_history, future):
        y_hat, theta, adj_knn, coeff = model(history_data=history, long_history_data=long_history, future_data=None,
                                             batch_seen=0, epoch=1)
        loss = step_loss(y_hat[..., [0]], future[..., [0]], theta, adj_knn, coeff, null_val=0.0)
        reducer.zero()
        loss.backward()
        reducer.reduce()                          # NCCL all-reduce (big tensor in place + one packed buffer) when world > 1
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    # ---- warm-up, then the device-resident measurement ----
    for i in range(max(args.warmup, 3)):
        train_step(*resident[i % 2])
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # kernels of libstep_b200.so enqueued inside the timed region, counted by the library itself (every launch site
    # goes through its check_launch); torch's own glue kernels are not included
    from step_b200 import lib as _lib
    k0 = int(_lib.load().step_launch_count())
    ms_res = timed(lambda i: train_step(*resident[i % 2]), args.steps)
    launches = int(_lib.load().step_launch_count()) - k0
    if args.only_resident:
        if rank == 0:
            sampler.stop()
            print(json.dumps({"only_resident": True, "ms_per_step": ms_res / args.steps, "gpu_launches": launches}))
        return

    # ---- end-to-end: pinned host batch -> H2D copies -> step -> loss read back, all inside the timed region.
    # The copies of step i+1 are issued on a side stream while step i computes (double-buffered device
    # staging), the way a prefetching data loader feeds the runner; every byte is still copied inside the
    # timed region and every step's loss is read back to the host.
    losses = []
    copy_stream = torch.cuda.Stream(dev)
    staging = [tuple(torch.empty_like(t, device=dev) for t in host[0]) for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def issue_copy(i):
        slot = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])            # the step that last used this slot is done with it
            for dst, src in zip(staging[slot], host[i % n_host]):
                dst.copy_(src, non_blocking=True)
            ready[slot].record(copy_stream)

    def e2e_step(i):
        if i == 0:
            issue_copy(0)
        issue_copy(i + 1)                                       # prefetch the next batch during this step
        slot = i % 2
        torch.cuda.current_stream(dev).wait_event(ready[slot])
        loss = train_step(*staging[slot])
        consumed[slot].record(torch.cuda.current_stream(dev))
        losses.append(loss.item())                              # D2H read of the step's result
    for ev in consumed:
        ev.record(torch.cuda.current_stream(dev))
    e2e_step(0)
    torch.cuda.synchronize(dev)
    for ev in consumed:
        ev.record(torch.cuda.current_stream(dev))
    ms_e2e = timed(e2e_step, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # ---- rooflines: kernels timed alone with CUDA events on the launching stream (after warm-up) ----
    pk = peaks()
    S_seq = B * NODES
    tokens = S_seq * PATCHES
    enc_flops = B * NODES * (4 * (PATCHES * (2 * 96 * 288 + 2 * 96 * 96 + 4 * 96 * 384) + 4 * 4 * PATCHES * PATCHES * 24)
                             + 2 * PATCHES * 12 * 96)
    lh = resident[0][1]

    def time_ms(fn, reps=10, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / reps

    def enc_once():
        with torch.no_grad():
            return model.tsformer(lh[..., [0]])
    enc_ms = time_ms(enc_once)
    enc_tflops = enc_flops / (enc_ms * 1e-3) / 1e12
    drop = 0.0 if args.no_dropout else 0.1
    roofline_other = [{"kernel": "TSFormer encoder, 21 launches (tc_embed + 4 x [QKV, attention, out+LN1, FFN1, FFN2+LN2])"
                       if args.precision == "bf16" else "TSFormer encoder, fp32 CUDA-core kernels",
                       "bound": "tensor", "achieved": enc_tflops, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                       "frac": enc_tflops / pk["bf16_tflops"], "ms": enc_ms, "useful_flops": enc_flops}]
    if args.precision == "bf16":
        L = ops._L()
        x_img = ops.tc_rows_to_image(torch.randn(tokens, 96, device=dev))
        w_in = ops.tc_pack_weight(torch.randn(288, 96, device=dev) * 0.15)
        b_in = torch.zeros(288, device=dev)
        q = torch.empty(L.step_tc_attn_image_bytes(S_seq, PATCHES, 0), device=dev, dtype=torch.uint8)
        k = torch.empty(L.step_tc_attn_image_bytes(S_seq, PATCHES, 1), device=dev, dtype=torch.uint8)
        v = torch.empty(L.step_tc_attn_image_bytes(S_seq, PATCHES, 1), device=dev, dtype=torch.uint8)
        o = torch.empty(((tokens + 127) // 128) * 96 * 256, device=dev, dtype=torch.uint8)
        st = ops._enter(x_img)
        ops.check(L.step_tc_qkv(x_img.data_ptr(), w_in.data_ptr(), b_in.data_ptr(), S_seq, PATCHES, q.data_ptr(), k.data_ptr(),
                                v.data_ptr(), st), "step_tc_qkv")
        att_ms = time_ms(lambda: ops.check(L.step_tc_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), S_seq,
                                                              PATCHES, drop, 1, st), "step_tc_attention"))
        att_flops = 4.0 * S_seq * 4 * PATCHES * PATCHES * 24            # useful (unpadded) QK^T + PV flops of one layer
        att_tflops = att_flops / (att_ms * 1e-3) / 1e12
        w1 = ops.tc_pack_weight(torch.randn(384, 96, device=dev) * 0.1)
        b1 = torch.zeros(384, device=dev)
        ffn_ms = time_ms(lambda: ops.tc_linear(x_img, w1, b1, tokens, 96, 384, 1))
        ffn_bytes = tokens * (96 + 384) * 2.0                            # bf16 activations in + out (weights stay in smem)
        ffn_gbs = ffn_bytes / (ffn_ms * 1e-3) / 1e9
        roofline = {"kernel": "tc_attn_kernel<%d,%d> (one TSFormer layer: S=QK^T, softmax, PV on tcgen05; %d sequences x 4 heads, "
                              "P=%d, head dim 24)" % (PATCHES, 1 if drop > 0 else 0, S_seq, PATCHES),
                    "bound": "tensor", "achieved": att_tflops, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                    "frac": att_tflops / pk["bf16_tflops"], "ms": att_ms, "useful_flops": att_flops,
                    "traffic": 968.8e6 if DATASET == "METR-LA" and B == 32 else None,
                    "traffic_source": "profiles/r01_ncu_full_v5_attn.txt (dram read + write per launch; algorithmic q,k,v,o "
                                      "bf16 bytes = 855 MB, the rest is row-tile / key padding of the operand images)",
                    "peak_source": pk["source"] + " (burst cuBLAS bf16, kernel timed alone)",
                    "note": "head dim 24 makes this kernel exp/issue-bound, not MMA-bound (SURVEY section 7): XU pipe 31%, issue slots "
                            "47%, tensor pipe 11% in the ncu capture; ~29% of the softmax warps' time is waiting on the MMA "
                            "round trip (one S accumulator per group fits in TMEM)"}
        roofline_other.append({"kernel": "tc_linear_kernel<1,0> (FFN1 [T,96]x[96,384] + bias + ReLU -> bf16 image)", "bound": "hbm",
                               "achieved": ffn_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ffn_gbs / pk["hbm_gbs"],
                               "ms": ffn_ms, "algorithmic_bytes": ffn_bytes})
        del x_img, q, k, v, o
    else:
        roofline = dict(roofline_other[0])
        roofline["traffic"] = None
        roofline["peak_source"] = pk["source"]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total_samples = B * world * args.steps
    value = total_samples / (ms_res * 1e-3)
    e2e = total_samples / (ms_e2e * 1e-3)
    out = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "fp32", "data": "synthetic",
        "config": {"workload": "STEP_%s fwd+loss+bwd, N=%d, per-GPU batch %d, P=%d (%d-step history), 12->12"
                               % (DATASET, NODES, B, PATCHES, PATCHES * 12),
                   "parallelism": "dp%d (batch-parallel, NCCL grad all-reduce)" % world if world > 1 else "single GPU",
                   "dropout": "off" if args.no_dropout else "live (TSFormer 0.1 in train(), gcn 0.3) as the reference trains",
                   "l2": "inputs > L2: each step streams a fresh 160 MB long-history batch and ~1 GB of activations",
                   "precision": ("TSFormer encoder bf16 operands / fp32 accumulate on tcgen05; graph learning + GWNet fp32"
                                 if args.precision == "bf16" else "fp32 everywhere"),
                   "ts_chunk_seqs": args.chunk_seqs, "weights": ("real TSFormer_METR-LA encoder" if DATASET == "METR-LA" else "seeded random TSFormer encoder")
                              + ", seeded random GWNet/DGL"},
        "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline, "roofline_other": roofline_other,
        "loss": losses[-1] if losses else None,
    }
    if not args.no_cpu_baseline:
        v, info = cpu_reference_run(2, 1, 2)
        info["value"] = v
        info["unit"] = "samples/s"
        out["cpu_baseline"] = info
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
