"""Tensor-level wrappers and autograd glue over the C ABI (include/step_b200.h).

PyTorch is used here only for device memory, streams and autograd bookkeeping; every
numeric hot op below is a hand-written sm_100a kernel reached through ctypes.  No op in this
file has a PyTorch/CPU fallback: a CPU tensor or a missing library raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch

from . import lib as _lib
from .lib import GwLayerGrads, GwLayerParams, TsLayerWeights, check

Tensor = torch.Tensor

# counters the benchmark reads: how many of OUR kernels were launched (host-side count)
launch_counter = {"kernels": 0}


def _L():
    return _lib.load()


def _f32(t: Tensor, name: str) -> Tensor:
    if not t.is_cuda:
        raise _lib.StepB200Error(f"{name}: expected a CUDA tensor - step_b200 has no CPU path")
    if t.dtype != torch.float32:
        raise _lib.StepB200Error(f"{name}: expected float32, got {t.dtype}")
    return t.contiguous()


def _p(t: Optional[Tensor]):
    return None if t is None else t.data_ptr()


def _enter(t: Tensor):
    """Select the tensor's device in the library's CUDA runtime and return the current torch stream."""
    check(_L().step_set_device(t.device.index), "step_set_device")
    return torch.cuda.current_stream(t.device).cuda_stream


# --------------------------------------------------------------------------- #
# TSFormer encoder
# --------------------------------------------------------------------------- #
def ts_layer_struct(layers: Sequence[Dict[str, Tensor]]):
    arr = (TsLayerWeights * len(layers))()
    keep = []
    for i, lw in enumerate(layers):
        for name in ("in_proj_w", "in_proj_b", "out_proj_w", "out_proj_b", "lin1_w", "lin1_b", "lin2_w", "lin2_b",
                     "norm1_w", "norm1_b", "norm2_w", "norm2_b"):
            t = _f32(lw[name], name)
            keep.append(t)
            setattr(arr[i], name, t.data_ptr())
    return arr, keep


def ts_encoder_forward(series: Tensor, patch_w: Tensor, patch_b: Tensor, pos: Tensor, layers: Sequence[Dict[str, Tensor]],
                       norm_w: Tensor, norm_b: Tensor, drop_p: float = 0.0, seed: int = 0, chunk_seqs: int = 0) -> Tensor:
    """series: [B, P*12, N] float32 view (any strides) -> hidden [B, N, P, 96]."""
    if not series.is_cuda or series.dtype != torch.float32:
        raise _lib.StepB200Error("ts_encoder_forward: series must be a float32 CUDA tensor")
    B, T, N = series.shape
    if T % 12 != 0:
        raise _lib.StepB200Error(f"ts_encoder_forward: history length {T} is not a multiple of the patch size 12")
    P = T // 12
    st = _enter(series)
    hidden = torch.empty(B, N, P, 96, device=series.device, dtype=torch.float32)
    S = B * N
    chunk = S if chunk_seqs <= 0 else min(chunk_seqs, S)
    ws_bytes = _L().step_ts_encoder_workspace_bytes(chunk, P)
    ws = torch.empty(ws_bytes, device=series.device, dtype=torch.uint8)
    arr, keep = ts_layer_struct(layers)
    pw, pb, ps = _f32(patch_w.reshape(96, 12), "patch_w"), _f32(patch_b, "patch_b"), _f32(pos, "pos")
    nw, nb = _f32(norm_w, "norm_w"), _f32(norm_b, "norm_b")
    sB, sT, sN = series.stride()
    check(_L().step_ts_encoder_fwd(series.data_ptr(), sB, sT, sN, B, N, P, pw.data_ptr(), pb.data_ptr(), ps.data_ptr(),
                                   arr, len(layers), nw.data_ptr(), nb.data_ptr(), hidden.data_ptr(), ws.data_ptr(),
                                   ws_bytes, chunk, float(drop_p), int(seed) & (2**64 - 1), st), "step_ts_encoder_fwd")
    n_chunks = (S + chunk - 1) // chunk
    launch_counter["kernels"] += 1 + n_chunks * len(layers) * 5
    return hidden


def ts_embed(series: Tensor, patch_w: Tensor, patch_b: Tensor, pos: Tensor, drop_p: float = 0.0, seed: int = 0) -> Tensor:
    """Patch + positional embedding (x sqrt(96), positional dropout): series [B, P*12, N] view -> tokens [B*N*P, 96]."""
    if not series.is_cuda or series.dtype != torch.float32:
        raise _lib.StepB200Error("ts_embed: series must be a float32 CUDA tensor")
    B, T, N = series.shape
    if T % 12 != 0:
        raise _lib.StepB200Error(f"ts_embed: history length {T} is not a multiple of the patch size 12")
    P = T // 12
    st = _enter(series)
    x = torch.empty(B * N * P, 96, device=series.device, dtype=torch.float32)
    pw, pb, ps = _f32(patch_w.reshape(96, 12), "patch_w"), _f32(patch_b, "patch_b"), _f32(pos, "pos")
    sB, sT, sN = series.stride()
    check(_L().step_ts_embed_fwd(series.data_ptr(), sB, sT, sN, B, N, P, pw.data_ptr(), pb.data_ptr(), ps.data_ptr(),
                                 x.data_ptr(), float(drop_p), int(seed) & (2**64 - 1), st), "step_ts_embed_fwd")
    launch_counter["kernels"] += 1
    return x


def ts_layers(x: Tensor, S: int, P: int, layers: Sequence[Dict[str, Tensor]], norm_w: Optional[Tensor] = None,
              norm_b: Optional[Tensor] = None, drop_p: float = 0.0, seed: int = 0) -> Tensor:
    """Transformer layer stack (+ optional final LayerNorm) on tokens x [S*P, 96] (already x sqrt(96)); returns a new tensor."""
    x = _f32(x, "x").clone()
    if x.shape != (S * P, 96):
        raise _lib.StepB200Error(f"ts_layers: x must be [{S * P}, 96], got {tuple(x.shape)}")
    st = _enter(x)
    ws_bytes = _L().step_ts_encoder_workspace_bytes(S, P)
    ws = torch.empty(ws_bytes, device=x.device, dtype=torch.uint8)
    arr, keep = ts_layer_struct(layers)
    nw = _f32(norm_w, "norm_w") if norm_w is not None else None
    nb = _f32(norm_b, "norm_b") if norm_b is not None else None
    check(_L().step_ts_layers_fwd(x.data_ptr(), S, P, arr, len(layers), _p(nw), _p(nb), ws.data_ptr(), ws_bytes,
                                  float(drop_p), int(seed) & (2**64 - 1), st), "step_ts_layers_fwd")
    launch_counter["kernels"] += len(layers) * 5
    return x


def linear(a: Tensor, w: Tensor, bias: Optional[Tensor], epilogue: int = 0, residual: Optional[Tensor] = None,
           ln_w: Optional[Tensor] = None, ln_b: Optional[Tensor] = None, drop_p: float = 0.0, seed: int = 0) -> Tensor:
    a = _f32(a, "a"); w = _f32(w, "w")
    M, K = a.shape
    Nout = w.shape[0]
    st = _enter(a)
    c = torch.empty(M, Nout, device=a.device, dtype=torch.float32)
    check(_L().step_linear_f32(a.data_ptr(), w.data_ptr(), _p(bias), c.data_ptr(), M, K, Nout, epilogue, _p(residual),
                               _p(ln_w), _p(ln_b), float(drop_p), int(seed) & (2**64 - 1), 7, st), "step_linear_f32")
    launch_counter["kernels"] += 1
    return c


def attention(qkv: Tensor, S: int, P: int, drop_p: float = 0.0, seed: int = 0) -> Tensor:
    qkv = _f32(qkv, "qkv")
    st = _enter(qkv)
    out = torch.empty(S * P, 96, device=qkv.device, dtype=torch.float32)
    check(_L().step_attn_fwd_f32(qkv.data_ptr(), out.data_ptr(), S, P, float(drop_p), int(seed), 0, st), "step_attn_fwd_f32")
    launch_counter["kernels"] += 1
    return out


# --------------------------------------------------------------------------- #
# kNN prior
# --------------------------------------------------------------------------- #
def cosine_gram(x: Tensor) -> Tensor:
    x = _f32(x, "x")
    B, N, D = x.shape
    st = _enter(x)
    norms = torch.empty(B, N, device=x.device, dtype=torch.float32)
    sim = torch.empty(B, N, N, device=x.device, dtype=torch.float32)
    check(_L().step_cosine_gram_f32(x.data_ptr(), B, N, D, norms.data_ptr(), sim.data_ptr(), st), "step_cosine_gram_f32")
    launch_counter["kernels"] += 2
    return sim


def topk_mask(sim: Tensor, k: int) -> Tensor:
    sim = _f32(sim, "sim")
    B, N, _ = sim.shape
    st = _enter(sim)
    adj = torch.empty_like(sim)
    check(_L().step_topk_mask_f32(sim.data_ptr(), B, N, int(k), adj.data_ptr(), st), "step_topk_mask_f32")
    launch_counter["kernels"] += 1
    return adj


def knn_prior(hidden: Tensor, k_total: int) -> Tensor:
    """hidden [B,N,P,d] -> adj_knn [B,N,N] (no gradient: discrete_graph_learning.py:108-110)."""
    B, N = hidden.shape[:2]
    with torch.no_grad():
        return topk_mask(cosine_gram(hidden.reshape(B, N, -1)), k_total)


# --------------------------------------------------------------------------- #
# edge logits + Gumbel sample
# --------------------------------------------------------------------------- #
class EdgeLogits(torch.autograd.Function):
    """(ut [F,N], v [N,F], cat_w [2,F], cat_b [2]) -> (logits [N,N,2], theta [N,N])."""

    @staticmethod
    def forward(ctx, ut, v, cat_w, cat_b):
        ut, v, cat_w, cat_b = _f32(ut, "ut"), _f32(v, "v"), _f32(cat_w, "cat_w"), _f32(cat_b, "cat_b")
        F_, N = ut.shape
        st = _enter(ut)
        logits = torch.empty(N, N, 2, device=ut.device, dtype=torch.float32)
        theta = torch.empty(N, N, device=ut.device, dtype=torch.float32)
        check(_L().step_edge_logits_fwd(ut.data_ptr(), v.data_ptr(), cat_w.data_ptr(), cat_b.data_ptr(), N, F_,
                                        logits.data_ptr(), theta.data_ptr(), st), "step_edge_logits_fwd")
        launch_counter["kernels"] += 1
        ctx.save_for_backward(ut, v, cat_w, theta)
        return logits, theta

    @staticmethod
    def backward(ctx, dlogits, dtheta):
        ut, v, cat_w, theta = ctx.saved_tensors
        F_, N = ut.shape
        dl = torch.zeros(N, N, 2, device=ut.device, dtype=torch.float32) if dlogits is None else dlogits.contiguous().clone()
        if dtheta is not None:
            t = dtheta * theta * (1.0 - theta)       # d softmax_0 / d(l0 - l1)
            dl[..., 0] += t
            dl[..., 1] -= t
        st = _enter(ut)
        dut = torch.empty_like(ut); dv = torch.empty_like(v)
        dcw = torch.empty_like(cat_w); dcb = torch.empty(2, device=ut.device, dtype=torch.float32)
        check(_L().step_edge_logits_bwd(dl.data_ptr(), ut.data_ptr(), v.data_ptr(), cat_w.data_ptr(), N, F_, dut.data_ptr(),
                                        dv.data_ptr(), dcw.data_ptr(), dcb.data_ptr(), st), "step_edge_logits_bwd")
        launch_counter["kernels"] += 2
        return dut, dv, dcw, dcb


class GumbelSample(torch.autograd.Function):
    """logits [N,N,2] -> sampled adjacency [B,N,N] in {0,1} with straight-through gradient."""

    @staticmethod
    def forward(ctx, logits, uniform, B, tau, seed):
        logits = _f32(logits, "logits")
        N = logits.shape[0]
        if uniform is not None:
            uniform = _f32(uniform, "uniform")
            if uniform.numel() != B * N * N * 2:
                raise _lib.StepB200Error("gumbel_sample: uniform must hold B*N*N*2 values")
        st = _enter(logits)
        sampled = torch.empty(B, N, N, device=logits.device, dtype=torch.float32)
        y0 = torch.empty(B, N, N, device=logits.device, dtype=torch.float32)
        check(_L().step_gumbel_sample_fwd(logits.data_ptr(), _p(uniform), B, N, float(tau), int(seed) & (2**64 - 1),
                                          sampled.data_ptr(), y0.data_ptr(), st), "step_gumbel_sample_fwd")
        launch_counter["kernels"] += 1
        ctx.save_for_backward(y0)
        ctx.tau = float(tau)
        return sampled

    @staticmethod
    def backward(ctx, dsampled):
        (y0,) = ctx.saved_tensors
        B, N, _ = y0.shape
        dsampled = _f32(dsampled, "dsampled")
        st = _enter(y0)
        dl = torch.empty(N, N, 2, device=y0.device, dtype=torch.float32)
        check(_L().step_gumbel_sample_bwd(dsampled.data_ptr(), y0.data_ptr(), B, N, ctx.tau, 0, dl.data_ptr(), st),
              "step_gumbel_sample_bwd")
        launch_counter["kernels"] += 1
        return dl, None, None, None, None


# --------------------------------------------------------------------------- #
# Graph WaveNet layer stack
# --------------------------------------------------------------------------- #
_GW_FIELDS = ("filter_w", "filter_b", "gate_w", "gate_b", "skip_w", "skip_b", "mlp_w", "mlp_b", "bn_w", "bn_b")


def _gw_time_extents(n_layers: int):
    t, outs = 13, []
    for i in range(n_layers):
        t -= 1 if i % 2 == 0 else 2
        outs.append(t)
    return outs


class GWNetStack(torch.autograd.Function):
    """x0 [B,13,N,32], P1/P2 [B,N,N], P3 [N,N], per-layer parameters -> (skip [B,N,256], bn_stats [L,4,32]).

    flat parameter order per layer: filter_w, filter_b, gate_w, gate_b, skip_w, skip_b, mlp_w, mlp_b, bn_w, bn_b
    (mlp_w/mlp_b/bn_w/bn_b = None for the last layer, whose gcn output is dead in the reference:
    step/step_arch/graphwavenet/model.py:217-218)."""

    @staticmethod
    def forward(ctx, x0, P1, P2, P3, training, drop_p, seed, bn_eval_stats, n_layers, *flat):
        x0, P1, P2, P3 = _f32(x0, "x0"), _f32(P1, "P1"), _f32(P2, "P2"), _f32(P3, "P3")
        B, T0, N, Cc = x0.shape
        if T0 != 13 or Cc != 32:
            raise _lib.StepB200Error(f"gwnet_stack: x0 must be [B,13,N,32], got {tuple(x0.shape)}")
        assert len(flat) == n_layers * 10
        params = [None if t is None else _f32(t, "gw param") for t in flat]
        st = _enter(x0)
        arr = (GwLayerParams * n_layers)()
        for i in range(n_layers):
            for j, name in enumerate(_GW_FIELDS):
                t = params[i * 10 + j]
                setattr(arr[i], name, None if t is None else t.data_ptr())
        skip = torch.empty(B, N, 256, device=x0.device, dtype=torch.float32)
        if training:
            bn_stats = torch.zeros(n_layers, 4, 32, device=x0.device, dtype=torch.float32)
        else:
            bn_stats = _f32(bn_eval_stats, "bn_eval_stats")
        stash = torch.empty(_L().step_gwnet_stash_floats(B, N, n_layers), device=x0.device, dtype=torch.float32)
        check(_L().step_gwnet_stack_fwd(x0.data_ptr(), P1.data_ptr(), P2.data_ptr(), P3.data_ptr(), arr, n_layers, B, N,
                                        1 if training else 0, float(drop_p), int(seed) & (2**64 - 1), skip.data_ptr(),
                                        bn_stats.data_ptr(), stash.data_ptr(), st), "step_gwnet_stack_fwd")
        n_gcn = sum(1 for i in range(n_layers) if params[i * 10 + 6] is not None)
        launch_counter["kernels"] += n_layers + (n_gcn if training else 0)
        ctx.save_for_backward(x0, P1, P2, P3, bn_stats, stash, *[p for p in params if p is not None])
        ctx.mask = [p is not None for p in params]
        ctx.cfg = (training, float(drop_p) if training else 0.0, int(seed) & (2**64 - 1), n_layers)
        ctx.mark_non_differentiable(bn_stats)
        return skip, bn_stats

    @staticmethod
    def backward(ctx, dskip, _dstats):
        training, drop_p, seed, n_layers = ctx.cfg
        if not training:
            raise _lib.StepB200Error("gwnet_stack: backward is only defined in training mode (batch statistics)")
        saved = list(ctx.saved_tensors)
        x0, P1, P2, P3, bn_stats, stash = saved[:6]
        it = iter(saved[6:])
        params = [next(it) if m else None for m in ctx.mask]
        B, _, N, _ = x0.shape
        dskip = _f32(dskip, "dskip")
        st = _enter(x0)
        P1t, P2t, P3t = P1.transpose(1, 2).contiguous(), P2.transpose(1, 2).contiguous(), P3.t().contiguous()
        arr = (GwLayerParams * n_layers)()
        garr = (GwLayerGrads * n_layers)()
        grads: List[Optional[Tensor]] = []
        for i in range(n_layers):
            for j, name in enumerate(_GW_FIELDS):
                t = params[i * 10 + j]
                setattr(arr[i], name, None if t is None else t.data_ptr())
                g = None if t is None else torch.empty_like(t)
                grads.append(g)
                setattr(garr[i], name, None if g is None else g.data_ptr())
        dx0 = torch.empty_like(x0)
        dP1, dP2, dP3 = torch.empty_like(P1), torch.empty_like(P2), torch.empty_like(P3)
        check(_L().step_gwnet_stack_bwd(dskip.data_ptr(), x0.data_ptr(), P1.data_ptr(), P2.data_ptr(), P3.data_ptr(),
                                        P1t.data_ptr(), P2t.data_ptr(), P3t.data_ptr(), arr, garr, n_layers, B, N,
                                        drop_p, seed, bn_stats.data_ptr(), stash.data_ptr(), dx0.data_ptr(),
                                        dP1.data_ptr(), dP2.data_ptr(), dP3.data_ptr(), st), "step_gwnet_stack_bwd")
        n_gcn = sum(1 for i in range(n_layers) if params[i * 10 + 6] is not None)
        launch_counter["kernels"] += 2 * n_layers + n_gcn + (n_layers - 1)
        return (dx0, dP1, dP2, dP3, None, None, None, None, None, *grads)


# --------------------------------------------------------------------------- #
# bf16 tensor-core encoder (tcgen05 / TMEM / TMA bulk)
# --------------------------------------------------------------------------- #
def tc_pack_weight(w: Tensor) -> Tensor:
    """fp32 [Nout, K] -> bf16 weight image (uint8 tensor of Nout*K*2 bytes)."""
    w = _f32(w, "w")
    Nout, K = w.shape
    st = _enter(w)
    img = torch.empty(Nout * K * 2, device=w.device, dtype=torch.uint8)
    check(_L().step_tc_pack_weight(w.data_ptr(), Nout, K, img.data_ptr(), st), "step_tc_pack_weight")
    launch_counter["kernels"] += 1
    return img


def tc_rows_to_image(x: Tensor) -> Tensor:
    x = _f32(x, "x")
    T, K = x.shape
    st = _enter(x)
    img = torch.empty(((T + 127) // 128) * K * 256, device=x.device, dtype=torch.uint8)
    check(_L().step_tc_rows_to_image(x.data_ptr(), T, K, img.data_ptr(), st), "step_tc_rows_to_image")
    return img


def tc_image_to_rows(img: Tensor, T: int, K: int) -> Tensor:
    st = _enter(img)
    x = torch.empty(T, K, device=img.device, dtype=torch.float32)
    check(_L().step_tc_image_to_rows(img.data_ptr(), T, K, x.data_ptr(), st), "step_tc_image_to_rows")
    return x


def tc_linear(a_img: Tensor, w_img: Tensor, bias: Tensor, T: int, K: int, Nout: int, mode: int, res_img: Optional[Tensor] = None,
              ln_w: Optional[Tensor] = None, ln_b: Optional[Tensor] = None, want_f32: bool = False):
    """mode 0 -> fp32 [T,Nout]; mode 1 -> ReLU image; mode 2 -> residual+LayerNorm image (and fp32 rows if want_f32)."""
    st = _enter(a_img)
    MT = (T + 127) // 128
    out_img = out_f32 = None
    if mode == 0 or want_f32:
        out_f32 = torch.empty(T, Nout, device=a_img.device, dtype=torch.float32)
    if mode in (1, 2):
        out_img = torch.empty(MT * Nout * 256, device=a_img.device, dtype=torch.uint8)
    check(_L().step_tc_linear(a_img.data_ptr(), w_img.data_ptr(), _f32(bias, "bias").data_ptr(), T, K, Nout, mode, _p(res_img),
                              _p(ln_w), _p(ln_b), _p(out_img), _p(out_f32), st), "step_tc_linear")
    launch_counter["kernels"] += 1
    return out_img, out_f32


def tc_qkv_attention(x_img: Tensor, w_img: Tensor, bias: Tensor, S: int, P: int, drop_p: float = 0.0, seed: int = 0) -> Tensor:
    """X image [S*P,96] -> O image [S*P,96] (QKV projection + softmax(QK^T)V, both on tcgen05)."""
    st = _enter(x_img)
    dev = x_img.device
    q = torch.empty(_L().step_tc_attn_image_bytes(S, P, 0), device=dev, dtype=torch.uint8)
    k = torch.empty(_L().step_tc_attn_image_bytes(S, P, 1), device=dev, dtype=torch.uint8)
    v = torch.empty(_L().step_tc_attn_image_bytes(S, P, 1), device=dev, dtype=torch.uint8)
    o = torch.empty(((S * P + 127) // 128) * 96 * 256, device=dev, dtype=torch.uint8)
    check(_L().step_tc_qkv(x_img.data_ptr(), w_img.data_ptr(), _f32(bias, "bias").data_ptr(), S, P, q.data_ptr(), k.data_ptr(),
                           v.data_ptr(), st), "step_tc_qkv")
    check(_L().step_tc_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), S, P, float(drop_p), int(seed), st),
          "step_tc_attention")
    launch_counter["kernels"] += 2
    return o


def ts_pack_layer_images(layers: Sequence[Dict[str, Tensor]]):
    """Pack the four weight matrices of every encoder layer into bf16 UMMA images (done once: the TSFormer is frozen)."""
    return [{"in_proj": tc_pack_weight(lw["in_proj_w"]), "out_proj": tc_pack_weight(lw["out_proj_w"]),
             "lin1": tc_pack_weight(lw["lin1_w"]), "lin2": tc_pack_weight(lw["lin2_w"])} for lw in layers]


def ts_encoder_forward_bf16(series: Tensor, patch_w: Tensor, patch_b: Tensor, pos: Tensor, layers: Sequence[Dict[str, Tensor]],
                            images, norm_w: Tensor, norm_b: Tensor, drop_p: float = 0.0, seed: int = 0, want_seq_image: bool = False):
    """series [B, P*12, N] view -> hidden [B, N, P, 96] fp32, computed on the tensor cores in bf16.
    want_seq_image: also return the bf16 sequence-major image of the hidden states (Gram operand)."""
    if not series.is_cuda or series.dtype != torch.float32:
        raise _lib.StepB200Error("ts_encoder_forward_bf16: series must be a float32 CUDA tensor")
    B, T, N = series.shape
    if T % 12 != 0:
        raise _lib.StepB200Error(f"ts_encoder_forward_bf16: history length {T} is not a multiple of the patch size 12")
    P = T // 12
    st = _enter(series)
    hidden = torch.empty(B, N, P, 96, device=series.device, dtype=torch.float32)
    ws_bytes = _L().step_ts_encoder_bf16_workspace_bytes(B, N, P)
    ws = torch.empty(ws_bytes, device=series.device, dtype=torch.uint8)
    seq_img = None
    if want_seq_image and (P * 12) % 8 == 0:
        seq_img = torch.empty(_L().step_tc_seq_image_bytes(B, N, P), device=series.device, dtype=torch.uint8)
    arr, keep = ts_layer_struct(layers)
    iarr = (_lib.TsLayerImages * len(layers))()
    for i, im in enumerate(images):
        for name in ("in_proj", "out_proj", "lin1", "lin2"):
            setattr(iarr[i], name, im[name].data_ptr())
    pw, pb, ps = _f32(patch_w.reshape(96, 12), "patch_w"), _f32(patch_b, "patch_b"), _f32(pos, "pos")
    nw, nb = _f32(norm_w, "norm_w"), _f32(norm_b, "norm_b")
    sB, sT, sN = series.stride()
    check(_L().step_ts_encoder_fwd_bf16(series.data_ptr(), sB, sT, sN, B, N, P, pw.data_ptr(), pb.data_ptr(), ps.data_ptr(), arr,
                                        iarr, len(layers), nw.data_ptr(), nb.data_ptr(), hidden.data_ptr(), _p(seq_img),
                                        ws.data_ptr(), ws_bytes, float(drop_p), int(seed) & (2**64 - 1), st),
          "step_ts_encoder_fwd_bf16")
    launch_counter["kernels"] += 1 + len(layers) * 5
    return (hidden, seq_img) if want_seq_image else hidden


def tc_hidden_to_seq_image(hidden: Tensor) -> Tensor:
    """fp32 hidden [B,N,P,96] -> bf16 sequence image (Gram operand)."""
    hidden = _f32(hidden, "hidden")
    B, N, P, _ = hidden.shape
    st = _enter(hidden)
    img = torch.empty(_L().step_tc_seq_image_bytes(B, N, P), device=hidden.device, dtype=torch.uint8)
    check(_L().step_tc_hidden_to_seq_image(hidden.data_ptr(), B, N, P, img.data_ptr(), st), "step_tc_hidden_to_seq_image")
    launch_counter["kernels"] += 1
    return img


def tc_cosine_gram(seq_img: Tensor, B: int, N: int, P: int) -> Tensor:
    """Cosine-similarity Gram matrix [B,N,N] from the encoder's bf16 sequence image (tcgen05)."""
    st = _enter(seq_img)
    scratch = torch.empty(B, N, N, device=seq_img.device, dtype=torch.float32)
    sim = torch.empty(B, N, N, device=seq_img.device, dtype=torch.float32)
    check(_L().step_tc_cosine_gram(seq_img.data_ptr(), B, N, P, scratch.data_ptr(), sim.data_ptr(), st), "step_tc_cosine_gram")
    launch_counter["kernels"] += 2
    return sim


# --------------------------------------------------------------------------- #
# discrete graph learning trunk (conv1 -> BN -> conv2 -> BN)
# --------------------------------------------------------------------------- #
class TrunkConv(torch.autograd.Function):
    """x [N, L0] -> (y2n [N, 16*(L0-18)], bn1_stats [4,8], bn2_stats [4,16]).  Parameters: conv1.w/b, bn1.w/b, conv2.w/b, bn2.w/b."""

    @staticmethod
    def forward(ctx, x, w1, b1, g1, be1, w2, b2, g2, be2, eps, training, eval_stats1, eval_stats2):
        x = _f32(x, "x")
        prm = [_f32(t, "trunk param") for t in (w1, b1, g1, be1, w2, b2, g2, be2)]
        N, L0 = x.shape
        L2 = L0 - 18
        st = _enter(x)
        dev = x.device
        if training:
            s1 = torch.empty(4, 8, device=dev, dtype=torch.float32)
            s2 = torch.empty(4, 16, device=dev, dtype=torch.float32)
        else:
            s1, s2 = _f32(eval_stats1, "eval_stats1"), _f32(eval_stats2, "eval_stats2")
        y2 = torch.empty(N, 16, L2, device=dev, dtype=torch.float32)
        y2n = torch.empty(N, 16, L2, device=dev, dtype=torch.float32)
        scratch = torch.empty(4096, device=dev, dtype=torch.uint8)
        check(_L().step_dgl_conv_fwd(x.data_ptr(), N, L0, *[t.data_ptr() for t in prm], float(eps), 1 if training else 0,
                                     s1.data_ptr(), s2.data_ptr(), y2.data_ptr(), y2n.data_ptr(), scratch.data_ptr(), st),
              "step_dgl_conv_fwd")
        launch_counter["kernels"] += 5 if training else 2
        ctx.save_for_backward(x, y2, s1, s2, *prm)
        ctx.eps, ctx.training = float(eps), bool(training)
        ctx.mark_non_differentiable(s1, s2)
        return y2n.view(N, 16 * L2), s1, s2

    @staticmethod
    def backward(ctx, dy2n, _d1, _d2):
        if not ctx.training:
            raise _lib.StepB200Error("TrunkConv: backward is only defined in training mode (batch statistics)")
        x, y2, s1, s2, w1, b1, g1, be1, w2, b2, g2, be2 = ctx.saved_tensors
        N, L0 = x.shape
        dy2n = _f32(dy2n, "dy2n")
        st = _enter(x)
        dev = x.device
        grads = [torch.empty_like(t) for t in (w1, b1, g1, be1, w2, b2, g2, be2)]
        dy1n = torch.empty(N, 8, L0 - 9, device=dev, dtype=torch.float32)
        scratch = torch.empty(4096, device=dev, dtype=torch.uint8)
        dw1, db1, dg1, dbe1, dw2, db2, dg2, dbe2 = grads
        check(_L().step_dgl_conv_bwd(dy2n.data_ptr(), x.data_ptr(), N, L0, w1.data_ptr(), b1.data_ptr(), g1.data_ptr(),
                                     w2.data_ptr(), g2.data_ptr(), ctx.eps, s1.data_ptr(), s2.data_ptr(), y2.data_ptr(),
                                     dy1n.data_ptr(), dw1.data_ptr(), db1.data_ptr(), dg1.data_ptr(), dbe1.data_ptr(),
                                     dw2.data_ptr(), db2.data_ptr(), dg2.data_ptr(), dbe2.data_ptr(), scratch.data_ptr(), st),
              "step_dgl_conv_bwd")
        launch_counter["kernels"] += 5
        return (None, *grads, None, None, None, None)


# --------------------------------------------------------------------------- #
# fused STEP loss (masked MAE + graph BCE), value and gradients from one pass
# --------------------------------------------------------------------------- #
class FusedStepLoss(torch.autograd.Function):
    """(pred [..], real [..] unscaled target feature, theta [N,N], adj_knn [B,N,N]) -> scalar loss."""

    @staticmethod
    def forward(ctx, pred, real, theta, adj_knn, coeff, null_val, mean, std):
        import math
        pred, real, theta, adj_knn = _f32(pred, "pred"), _f32(real, "real"), _f32(theta, "theta"), _f32(adj_knn, "adj_knn")
        B, N, _ = adj_knn.shape
        st = _enter(pred)
        dev = pred.device
        loss = torch.empty(1, device=dev, dtype=torch.float32)
        dpred = torch.empty_like(pred)
        dtheta = torch.empty_like(theta)
        scratch = torch.empty(64, device=dev, dtype=torch.uint8)
        nan_mask = 1 if (isinstance(null_val, float) and math.isnan(null_val)) else 0
        check(_L().step_loss_fwd_bwd(pred.data_ptr(), real.data_ptr(), pred.numel(), float(mean), float(std),
                                     0.0 if nan_mask else float(null_val), nan_mask, theta.data_ptr(), adj_knn.data_ptr(), B, N,
                                     float(coeff), loss.data_ptr(), dpred.data_ptr(), dtheta.data_ptr(), scratch.data_ptr(), st),
              "step_loss_fwd_bwd")
        launch_counter["kernels"] += 2
        ctx.save_for_backward(dpred, dtheta)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        dpred, dtheta = ctx.saved_tensors
        return dpred * g, None, dtheta * g, None, None, None, None, None
