#!/usr/bin/env python
"""A/B of two builds of the attention kernel on the SAME GPU: the library in the tree (B) against another shared object
(A, e.g. an earlier commit's tc_encoder.cu built into ab_tmp/libA.so).  Each library prepares its own Q/K/V images with its
own QKV epilogue; the attention launch is timed with CUDA events, alternating A and B.  Usage: attn_ab.py path/to/libA.so"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from step_b200 import ops  # noqa: E402


def bind(lib):
    vp = C.c_void_p
    lib.step_tc_attn_image_bytes.restype = C.c_size_t
    lib.step_tc_attn_image_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.step_tc_qkv.restype = C.c_int
    lib.step_tc_qkv.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp]
    lib.step_tc_attention.restype = C.c_int
    lib.step_tc_attention.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_float, C.c_ulonglong, vp]
    return lib


def main():
    libs = {"A": bind(C.CDLL(sys.argv[1])), "B": bind(C.CDLL(os.path.join(os.path.dirname(__file__), "..", "step_b200", "libstep_b200.so")))}
    S, P, drop = 6624, 168, 0.1
    dev = torch.device("cuda", 0)
    T = S * P
    x_img = ops.tc_rows_to_image(torch.randn(T, 96, device=dev))
    w = ops.tc_pack_weight(torch.randn(288, 96, device=dev) * 0.15)
    b = torch.zeros(288, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    bufs = {}
    for name, L in libs.items():
        q = torch.empty(L.step_tc_attn_image_bytes(S, P, 0), device=dev, dtype=torch.uint8)
        k = torch.empty(L.step_tc_attn_image_bytes(S, P, 1), device=dev, dtype=torch.uint8)
        v = torch.empty(L.step_tc_attn_image_bytes(S, P, 1), device=dev, dtype=torch.uint8)
        o = torch.empty(((T + 127) // 128) * 96 * 256, device=dev, dtype=torch.uint8)
        bound = torch.empty(L.step_tc_attn_image_bytes(S, P, 2) // 4, device=dev, dtype=torch.float32)
        assert L.step_tc_qkv(x_img.data_ptr(), w.data_ptr(), b.data_ptr(), S, P, q.data_ptr(), k.data_ptr(), v.data_ptr(), bound.data_ptr(), st) == 0
        bufs[name] = (q, k, v, o, bound)
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    times = {"A": [], "B": []}
    for rep in range(12):
        for name, L in libs.items():
            q, k, v, o, bound = bufs[name]
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            assert L.step_tc_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), bound.data_ptr(), S, P, drop, 1, st) == 0
            e1.record()
            torch.cuda.synchronize()
            if rep >= 2:
                times[name].append(e0.elapsed_time(e1))
    for name in times:
        t = sorted(times[name])
        print(f"{name}: median {t[len(t) // 2]:.4f} ms  min {t[0]:.4f}  max {t[-1]:.4f}  ({len(t)} launches, L2 flushed before each)")
    oa = ops.tc_image_to_rows(bufs["A"][3], T, 96)
    ob = ops.tc_image_to_rows(bufs["B"][3], T, 96)
    print("max |O_A - O_B| (different dropout streams expected):", float((oa - ob).abs().max()))


if __name__ == "__main__":
    main()
