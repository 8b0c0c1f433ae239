#!/usr/bin/env python
"""Timeline of tc_attn_kernel's CTA 0 from step_tc_attention_trace (clock64 stamps): per iteration and role the waits and
the work phases in SM cycles.  Usage: python tools/attn_trace.py [S] [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from step_b200 import ops  # noqa: E402


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 6624
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    P, drop = 168, 0.1
    dev = torch.device("cuda", 0)
    L = ops._L()
    T = S * P
    x_img = ops.tc_rows_to_image(torch.randn(T, 96, device=dev))
    w = ops.tc_pack_weight(torch.randn(288, 96, device=dev) * 0.15)
    b = torch.zeros(288, device=dev)
    q = torch.empty(L.step_tc_attn_image_bytes(S, P, 0), device=dev, dtype=torch.uint8)
    k = torch.empty(L.step_tc_attn_image_bytes(S, P, 1), device=dev, dtype=torch.uint8)
    v = torch.empty(L.step_tc_attn_image_bytes(S, P, 1), device=dev, dtype=torch.uint8)
    o = torch.empty(((T + 127) // 128) * 96 * 256, device=dev, dtype=torch.uint8)
    bound = torch.empty(L.step_tc_attn_image_bytes(S, P, 2) // 4, device=dev, dtype=torch.float32)
    st = ops._enter(x_img)
    ops.check(L.step_tc_qkv(x_img.data_ptr(), w.data_ptr(), b.data_ptr(), S, P, q.data_ptr(), k.data_ptr(), v.data_ptr(),
                            bound.data_ptr(), st), "qkv")
    trace = torch.zeros(iters, 3, 8, device=dev, dtype=torch.int64)
    for _ in range(2):
        ops.check(L.step_tc_attention_trace(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), bound.data_ptr(), S, P, drop, 1,
                                            trace.data_ptr(), iters, st), "trace")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.check(L.step_tc_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), bound.data_ptr(), S, P, drop, 1, st), "attn")
    e1.record()
    torch.cuda.synchronize()
    print(f"kernel time {e0.elapsed_time(e1) / 20:.4f} ms (20 launches, CUDA events, no trace)")
    t = trace.cpu()
    t0 = int(t[0, 2, 0])
    print("cycles relative to the MMA issuer's first stamp; iteration i: group i&1 (even = full 128-row tile, odd = 40-row tail)")
    print("  i | softmax group: waitS  S_ready  P_pub  O_ready  stored | issuer: enter  QKV_ok  Sslot_ok | PV(i): enter  P_ok")
    for i in range(iters):
        g = i & 1
        sg = [int(x) - t0 for x in t[i, g, :5]]
        mm = [int(x) - t0 for x in t[i, 2, :5]]
        print(f"{i:3d} | {sg[0]:7d} {sg[1]:7d} {sg[2]:7d} {sg[3]:7d} {sg[4]:7d} | {mm[0]:7d} {mm[1]:7d} {mm[2]:7d} | {mm[3]:7d} {mm[4]:7d}")
    for g in (0, 1):
        rows = [t[i, g] for i in range(8 + g, iters, 2)]
        import statistics as st_
        per = [int(rows[j + 1][0] - rows[j][0]) for j in range(len(rows) - 1)]
        sm = [int(r[2] - r[1]) for r in rows]
        ws = [int(r[1] - r[0]) for r in rows]
        wo = [int(r[3] - r[2]) for r in rows]
        ep = [int(r[4] - r[3]) for r in rows]
        print(f"group {g}: period {st_.median(per)}  wait_S {st_.median(ws)}  softmax {st_.median(sm)}  wait_O {st_.median(wo)}  epilogue {st_.median(ep)}")


if __name__ == "__main__":
    main()
