#!/usr/bin/env python
"""Micro-benchmarks of individual kernels at the STEP_METR-LA full size (CUDA events, L2 flushed between
repetitions by the working set itself: every operand set is > 126 MB).  Usage: python tools/kbench.py [names...]"""
import os
import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from step_b200 import ops  # noqa: E402

DEV = torch.device("cuda:0")
S, P = 32 * 207, 168
T = S * P


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def bench_tc_linear():
    x_img = ops.tc_rows_to_image(torch.randn(T, 96, device=DEV))
    h_img = torch.zeros(((T + 127) // 128) * 384 * 256, device=DEV, dtype=torch.uint8)
    lw, lb = torch.ones(96, device=DEV), torch.zeros(96, device=DEV)
    for name, K, Nout, mode, a in (("out+LN1 [96->96]", 96, 96, 2, x_img), ("FFN1 relu [96->384]", 96, 384, 1, x_img),
                                   ("FFN2+LN2 [384->96]", 384, 96, 2, h_img), ("f32out [96->288]", 96, 288, 0, x_img)):
        w = ops.tc_pack_weight(torch.randn(Nout, K, device=DEV) * 0.1)
        b = torch.zeros(Nout, device=DEV)
        ms = timeit(lambda: ops.tc_linear(a, w, b, T, K, Nout, mode, res_img=x_img if mode == 2 else None, ln_w=lw, ln_b=lb))
        print(f"tc_linear {name:22s} {ms * 1e3:8.1f} us   {2.0 * T * K * Nout / ms / 1e9:7.1f} TFLOP/s")


def bench_tc_attn():
    x_img = ops.tc_rows_to_image(torch.randn(T, 96, device=DEV))
    w = ops.tc_pack_weight(torch.randn(288, 96, device=DEV) * 0.15)
    b = torch.zeros(288, device=DEV)
    L = ops._L()
    q = torch.empty(L.step_tc_attn_image_bytes(S, P, 0), device=DEV, dtype=torch.uint8)
    k = torch.empty(L.step_tc_attn_image_bytes(S, P, 1), device=DEV, dtype=torch.uint8)
    v = torch.empty(L.step_tc_attn_image_bytes(S, P, 1), device=DEV, dtype=torch.uint8)
    o = torch.empty(((T + 127) // 128) * 96 * 256, device=DEV, dtype=torch.uint8)
    st = ops._enter(x_img)
    ms = timeit(lambda: ops.check(L.step_tc_qkv(x_img.data_ptr(), w.data_ptr(), b.data_ptr(), S, P, q.data_ptr(), k.data_ptr(),
                                                v.data_ptr(), st), "qkv"))
    print(f"tc_qkv                           {ms * 1e3:8.1f} us")
    for p in (0.0, 0.1):
        ms = timeit(lambda: ops.check(L.step_tc_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), S, P, p, 1, st), "attn"))
        print(f"tc_attention drop={p}             {ms * 1e3:8.1f} us   {4.0 * S * 4 * P * P * 24 / ms / 1e9:7.1f} useful TFLOP/s")


def bench_gram():
    x = torch.randn(32, 207, 16128, device=DEV)
    ms = timeit(lambda: ops.cosine_gram(x))
    print(f"cosine_gram fp32                 {ms * 1e3:8.1f} us   {2.0 * 32 * 207 * 207 * 16128 / ms / 1e9:7.1f} TFLOP/s")
    sim = ops.cosine_gram(x)
    ms = timeit(lambda: ops.topk_mask(sim, 2070))
    print(f"topk_mask                        {ms * 1e3:8.1f} us")


ALL = {"tc_linear": bench_tc_linear, "tc_attn": bench_tc_attn, "gram": bench_gram}
if __name__ == "__main__":
    for n in (sys.argv[1:] or list(ALL)):
        ALL[n]()
