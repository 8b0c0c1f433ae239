#!/usr/bin/env python
"""Debug aid for tc_ffn_kernel (library built with STEP_B200_NVCC_FLAGS=-DSTEP_FFN_DEBUG): runs the bf16 encoder with the fused
feed-forward kernel and, if the launch dies, prints at which wait site every warp of every CTA was blocked (host-mapped slots)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "tests")))


def main():
    from oracle import step_oracle as O
    from step_b200 import ops
    dev = torch.device("cuda", 0)
    B, N, P = 2, 9, 168
    slots = torch.zeros(148 * 16, dtype=torch.int32).pin_memory()
    os.environ["STEP_FFN_DEBUG_PTR"] = hex(slots.data_ptr())
    sd = O.synthetic_tsformer_params(2)
    layers = []
    for i in range(4):
        p = f"encoder.transformer_encoder.layers.{i}."
        layers.append({k: sd[p + v].to(dev) for k, v in dict(
            in_proj_w="self_attn.in_proj_weight", in_proj_b="self_attn.in_proj_bias", out_proj_w="self_attn.out_proj.weight",
            out_proj_b="self_attn.out_proj.bias", lin1_w="linear1.weight", lin1_b="linear1.bias", lin2_w="linear2.weight",
            lin2_b="linear2.bias", norm1_w="norm1.weight", norm1_b="norm1.bias", norm2_w="norm2.weight", norm2_b="norm2.bias").items()})
    images = ops.ts_pack_layer_images(layers)
    series = torch.randn(B, P * 12, N, device=dev)
    args = (series, sd["patch_embedding.input_embedding.weight"].to(dev), sd["patch_embedding.input_embedding.bias"].to(dev),
            sd["positional_encoding.position_embedding"].to(dev), layers, images, sd["encoder_norm.weight"].to(dev),
            sd["encoder_norm.bias"].to(dev))
    try:
        h, _ = ops.ts_encoder_forward_bf16(*args, drop_p=0.0, seed=21, want_seq_image=True)
        torch.cuda.synchronize()
        print("completed, finite:", bool(torch.isfinite(h).all()))
    except Exception as e:  # noqa: BLE001
        print("FAILED:", str(e).splitlines()[0])
    s = slots.view(148, 16)[:24, :10]
    print("rows = CTAs, columns = warps 0..9; value = wait site being blocked on (100 + site = passed)")
    print(s)


if __name__ == "__main__":
    main()
