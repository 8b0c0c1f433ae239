"""GPU tests of the tcgen05 (bf16 tensor-core) encoder kernels.

bf16 operands with fp32 accumulation: the comparison target for a single GEMM is the same product of
bf16-rounded operands in fp32 (tight), for attention / the whole encoder the fp32 oracle with a bf16-level
tolerance (stated in each test)."""
import math

import pytest
import torch

from oracle import step_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def bf(x):
    return x.to(torch.bfloat16).float()


def test_image_roundtrip():
    from step_b200 import ops
    x = torch.randn(300, 96)
    img = ops.tc_rows_to_image(x.to(DEV))
    back = ops.tc_image_to_rows(img, 300, 96).cpu()
    assert torch.equal(back, bf(x))


@pytest.mark.parametrize("T,K,Nout", [(128, 96, 96), (300, 96, 288), (1000, 96, 384), (515, 384, 96), (40000, 96, 288)])
def test_tc_linear_f32_out(T, K, Nout):
    from step_b200 import ops
    g = torch.Generator().manual_seed(T + K + Nout)
    x, w, b = torch.randn(T, K, generator=g), torch.randn(Nout, K, generator=g) * 0.1, torch.randn(Nout, generator=g)
    ref = bf(x) @ bf(w).t() + b
    _, out = ops.tc_linear(ops.tc_rows_to_image(x.to(DEV)), ops.tc_pack_weight(w.to(DEV)), b.to(DEV), T, K, Nout, 0)
    err = (out.cpu() - ref).abs().max().item()
    assert err < 1e-3 * max(1.0, ref.abs().max().item()), err


def test_tc_linear_relu_and_resln():
    from step_b200 import ops
    g = torch.Generator().manual_seed(3)
    T = 700
    x, w1, b1 = torch.randn(T, 96, generator=g), torch.randn(384, 96, generator=g) * 0.1, torch.randn(384, generator=g) * 0.1
    w2, b2 = torch.randn(96, 384, generator=g) * 0.1, torch.randn(96, generator=g) * 0.1
    lw, lb = torch.rand(96, generator=g) + 0.5, torch.randn(96, generator=g) * 0.1
    x_img = ops.tc_rows_to_image(x.to(DEV))
    h_img, _ = ops.tc_linear(x_img, ops.tc_pack_weight(w1.to(DEV)), b1.to(DEV), T, 96, 384, 1)
    h_ref = bf(torch.relu(bf(x) @ bf(w1).t() + b1))
    h = ops.tc_image_to_rows(h_img, T, 384).cpu()
    assert (h - h_ref).abs().max().item() < 0.02
    y_img, y = ops.tc_linear(h_img, ops.tc_pack_weight(w2.to(DEV)), b2.to(DEV), T, 384, 96, 2, res_img=x_img, ln_w=lw.to(DEV),
                             ln_b=lb.to(DEV), want_f32=True)
    y_ref = O._layer_norm(h @ bf(w2).t() + b2 + bf(x), lw, lb)
    assert (y.cpu() - y_ref).abs().max().item() < 2e-3
    assert (ops.tc_image_to_rows(y_img, T, 96).cpu() - bf(y.cpu())).abs().max().item() == 0.0


# P > 176: key-split kernel (two 176-key blocks merged through shared memory); 336 = PEMS03/04/08 histories
@pytest.mark.parametrize("S,P", [(3, 168), (16, 168), (5, 100), (2, 24), (3, 336), (301, 336), (4, 200), (2, 177), (3, 352)])
def test_tc_qkv_attention(S, P):
    from step_b200 import ops
    g = torch.Generator().manual_seed(S * 7 + P)
    T = S * P
    x = torch.randn(T, 96, generator=g)
    w, b = torch.randn(288, 96, generator=g) * 0.15, torch.randn(288, generator=g) * 0.1
    qkv = bf(x) @ bf(w).t() + b
    q, k, v = qkv.view(S, P, 288).split(96, -1)
    sh = lambda t: t.reshape(S, P, 4, 24).transpose(1, 2)
    att = torch.softmax(sh(q) @ sh(k).transpose(-1, -2) / math.sqrt(24), -1)
    ref = (att @ sh(v)).transpose(1, 2).reshape(T, 96)
    # the exact-row-maximum route is checked here; test_tc_attention_bounded_max_matches_exact_max covers the bounded one
    o_img = ops.tc_qkv_attention(ops.tc_rows_to_image(x.to(DEV)), ops.tc_pack_weight(w.to(DEV)), b.to(DEV), S, P, bounded_max=False)
    out = ops.tc_image_to_rows(o_img, T, 96).cpu()
    err = (out - ref).abs()
    # vs exact fp32 softmax: bf16 q/k/v/p carry ~2^-8 relative error per operand; values are O(1).  The max grows with
    # the number of outputs (a CPU emulation of the same roundings reproduces it), the mean does not.
    assert err.max().item() < 1e-1 and err.mean().item() < 4e-3, (err.max().item(), err.mean().item())
    # vs a CPU emulation of the kernel's own roundings (scaled q, k, v and p rounded to bf16, fp32 statistics,
    # normalisation by the unrounded row sum, bf16 output): only ex2.approx / summation order / rounding flips remain
    qs = bf(sh(q) * (1.0 / math.sqrt(24)) * 1.4426950408889634)
    sc = qs @ bf(sh(k)).transpose(-1, -2)
    p = torch.exp2(sc - sc.max(-1, keepdim=True).values)
    emu = bf((bf(p) @ bf(sh(v))) / p.sum(-1, keepdim=True)).transpose(1, 2).reshape(T, 96)
    e2 = (out - emu).abs()
    rel = (e2 / (emu.abs() + 1.0)).max().item()          # one bf16 ulp of the output is 2^-8 relative
    assert rel < 1e-2 and e2.mean().item() < 3e-4, (rel, e2.max().item(), e2.mean().item())


def _layers(sd):
    out = []
    for i in range(4):
        p = f"encoder.transformer_encoder.layers.{i}."
        out.append({"in_proj_w": sd[p + "self_attn.in_proj_weight"], "in_proj_b": sd[p + "self_attn.in_proj_bias"],
                    "out_proj_w": sd[p + "self_attn.out_proj.weight"], "out_proj_b": sd[p + "self_attn.out_proj.bias"],
                    "lin1_w": sd[p + "linear1.weight"], "lin1_b": sd[p + "linear1.bias"],
                    "lin2_w": sd[p + "linear2.weight"], "lin2_b": sd[p + "linear2.bias"],
                    "norm1_w": sd[p + "norm1.weight"], "norm1_b": sd[p + "norm1.bias"],
                    "norm2_w": sd[p + "norm2.weight"], "norm2_b": sd[p + "norm2.bias"]})
    return [{k: v.to(DEV) for k, v in l.items()} for l in out]


@pytest.mark.parametrize("B,N,P,real", [(2, 9, 168, False), (1, 40, 168, True), (3, 5, 24, False), (1, 11, 336, False)])
def test_ts_encoder_bf16_close_to_oracle(B, N, P, real):
    import os
    from conftest import GOLDEN
    from step_b200 import ops
    sd = torch.load(os.path.join(GOLDEN, "tsformer_METR-LA_state.pt")) if real else O.synthetic_tsformer_params(1)
    g = torch.Generator().manual_seed(5)
    long_history = torch.randn(B, P * 12, N, 3, generator=g)
    ref = O.tsformer_encode(sd, long_history[..., [0]])
    layers = _layers(sd)
    images = ops.ts_pack_layer_images(layers)
    out = ops.ts_encoder_forward_bf16(long_history.to(DEV)[..., 0], sd["patch_embedding.input_embedding.weight"].to(DEV),
                                      sd["patch_embedding.input_embedding.bias"].to(DEV),
                                      sd["positional_encoding.position_embedding"].to(DEV), layers, images,
                                      sd["encoder_norm.weight"].to(DEV), sd["encoder_norm.bias"].to(DEV)).cpu()
    err = (out - ref).abs()
    print("bf16 encoder vs fp32 oracle: max", err.max().item(), "mean", err.mean().item(), "ref absmean", ref.abs().mean().item())
    # 4 post-norm layers in bf16: stated tolerance = 3e-2 mean abs (hidden states are O(0.3-1))
    assert torch.isfinite(out).all()
    assert err.mean().item() < 3e-2 and err.max().item() < 0.5


@pytest.mark.parametrize("B,N,P", [(2, 50, 24), (2, 207, 168), (1, 300, 24)])
def test_tc_cosine_gram_from_encoder_image(B, N, P):
    from step_b200 import ops
    sd = O.synthetic_tsformer_params(2)
    g = torch.Generator().manual_seed(B + N + P)
    long_history = torch.randn(B, P * 12, N, 1, generator=g)
    layers = _layers(sd)
    images = ops.ts_pack_layer_images(layers)
    hidden, img = ops.ts_encoder_forward_bf16(long_history.to(DEV)[..., 0], sd["patch_embedding.input_embedding.weight"].to(DEV),
                                              sd["patch_embedding.input_embedding.bias"].to(DEV),
                                              sd["positional_encoding.position_embedding"].to(DEV), layers, images,
                                              sd["encoder_norm.weight"].to(DEV), sd["encoder_norm.bias"].to(DEV),
                                              want_seq_image=True)
    sim = ops.tc_cosine_gram(img, B, N, P).cpu()
    x = bf(hidden.cpu()).reshape(B, N, -1)                 # the image holds the bf16-rounded hidden states
    ref = O.cosine_similarity_gram(x.double()).float()
    assert (sim - ref).abs().max().item() < 2e-4
    assert (sim.diagonal(dim1=1, dim2=2) - 1).abs().max().item() < 1e-5


@pytest.mark.parametrize("S,P,scale", [(5, 168, 0.1), (3, 336, 0.1), (4, 100, 0.1), (3, 200, 0.1), (5, 168, 1.5), (2, 336, 1.5)])
def test_tc_attention_bounded_max_matches_exact_max(S, P, scale):
    """The one-pass softmax against the Cauchy-Schwarz bound |q_i| max_j |k_j| (attention kernel, bound <= 40) and the exact
    two-pass row maximum are the same softmax up to the bf16 rounding of the probabilities (a different power-of-two-free
    scale re-draws every rounding: the outputs differ by independent rounding noise, not by a bias); projection weights
    scaled x10 push the bound past 40 and exercise the warp-uniform fall-back to the exact route (then both runs are
    bit-identical)."""
    from step_b200 import ops
    g = torch.Generator().manual_seed(S + P)
    T = S * P
    x = torch.randn(T, 96, generator=g)
    w, b = torch.randn(288, 96, generator=g) * scale, torch.randn(288, generator=g) * 0.1
    x_img, w_img = ops.tc_rows_to_image(x.to(DEV)), ops.tc_pack_weight(w.to(DEV))
    got = [ops.tc_image_to_rows(ops.tc_qkv_attention(x_img, w_img, b.to(DEV), S, P, bounded_max=bm), T, 96).cpu() for bm in (True, False)]
    assert torch.isfinite(got[0]).all()
    qkv = bf(x) @ bf(w).t() + b
    q, k, v = qkv.view(S, P, 288).split(96, -1)
    sh = lambda t: t.reshape(S, P, 4, 24).transpose(1, 2)
    bound = (sh(q).norm(dim=-1) * sh(k).norm(dim=-1).amax(-1, keepdim=True)) * (1.4426950408889634 / math.sqrt(24))
    if scale > 1.0:
        assert bound.min().item() > 60                    # every warp holds a row past the limit -> exact route everywhere
        assert torch.equal(got[0], got[1])
        return
    assert bound.max().item() < 38                        # every row takes the bounded route
    assert not torch.equal(got[0], got[1])
    err = (got[0] - got[1]).abs()
    assert (err / (got[1].abs() + 1.0)).max().item() < 1e-2 and err.mean().item() < 1.5e-3
    assert abs((got[0] - got[1]).mean().item()) < 5e-5   # no bias between the two routes
    att = torch.softmax(sh(q) @ sh(k).transpose(-1, -2) / math.sqrt(24), -1)
    ref = (att @ sh(v)).transpose(1, 2).reshape(T, 96)
    e_b, e_x = (got[0] - ref).abs(), (got[1] - ref).abs()
    assert e_b.max().item() < 1e-1 and e_b.mean().item() < 4e-3
    assert e_b.mean().item() < 1.25 * e_x.mean().item() + 1e-4   # as close to the fp32 softmax as the exact route


def test_tc_attention_dropout_per_key_rates():
    """Keep rate of the attention-probability dropout per key residue class (both 16-bit halves of the random word and all
    positions of an 8-key chunk): v_j = one-hot(j mod 24), q = k = 0 -> O[row, head, d] = mean keep over keys j = d mod 24."""
    from step_b200 import ops
    S, P, p = 24, 168, 0.1
    T = S * P
    x = torch.zeros(T, 96)
    x[torch.arange(T), torch.arange(T) % P % 24] = 1.0
    w = torch.zeros(288, 96)
    for h in range(4):
        for d in range(24):
            w[192 + h * 24 + d, d] = 1.0
    o_img = ops.tc_qkv_attention(ops.tc_rows_to_image(x.to(DEV)), ops.tc_pack_weight(w.to(DEV)), torch.zeros(288, device=DEV), S, P,
                                 drop_p=p, seed=11)
    out = ops.tc_image_to_rows(o_img, T, 96).cpu().view(T, 4, 24) * (1 - p) * P / 7.0      # mean keep of the 7 keys of class d
    rate = out.mean((0, 1))
    n = T * 4 * 7
    assert (rate - (1 - p)).abs().max().item() < 5 * math.sqrt(p * (1 - p) / n) + 2e-3, rate
    # different keys of one row are dropped independently: variance of the per-row kept count is binomial
    cnt = out.sum(-1) * 7.0
    assert cnt.var().item() == pytest.approx(P * p * (1 - p), rel=0.1)


def test_tc_attention_dropout_statistics():
    """Counter-hash dropout of the tensor-core path: keep probability 1 - p, binomial spread, reproducible per seed."""
    from step_b200 import ops
    S, P, p = 16, 168, 0.1
    T = S * P
    x_img = ops.tc_rows_to_image(torch.zeros(T, 96, device=DEV))
    w = ops.tc_pack_weight(torch.zeros(288, 96, device=DEV))
    b = torch.zeros(288); b[192:] = 1.0                       # q = k = 0 (uniform attention), v = 1
    outs = []
    for seed in (5, 5, 6):
        o_img = ops.tc_qkv_attention(x_img, w, b.to(DEV), S, P, drop_p=p, seed=seed)
        outs.append(ops.tc_image_to_rows(o_img, T, 96).cpu())
    kept = outs[0] * (1 - p)                                   # = (#kept keys) / P per (row, head), replicated over 24 dims
    assert abs(kept.mean().item() - (1 - p)) < 3e-3
    assert kept[:, ::24].std().item() == pytest.approx(math.sqrt(p * (1 - p) / P), rel=0.2)
    assert torch.equal(outs[0], outs[1]) and not torch.equal(outs[0], outs[2])


@pytest.mark.parametrize("B,N,P,drop", [(2, 9, 168, 0.0), (1, 11, 336, 0.0), (3, 5, 24, 0.0), (2, 13, 168, 0.1)])
def test_fused_token_block_kernel_matches_separate_kernels(B, N, P, drop):
    """The optional fused kernel per layer (STEP_B200_TS_FUSED=1: out-proj + LN1 + FFN + LN2 + next layer's QKV,
    intermediates in shared memory) vs the default four separate token GEMM launches: same MMA order, same dropout counters -> the hidden states and the Gram operand image are
    bit-identical, with and without dropout."""
    from step_b200 import ops
    sd = O.synthetic_tsformer_params(2)
    g = torch.Generator().manual_seed(7)
    series = torch.randn(B, P * 12, N, generator=g).to(DEV)
    layers = _layers(sd)
    images = ops.ts_pack_layer_images(layers)
    args = (series, sd["patch_embedding.input_embedding.weight"].to(DEV), sd["patch_embedding.input_embedding.bias"].to(DEV),
            sd["positional_encoding.position_embedding"].to(DEV), layers, images, sd["encoder_norm.weight"].to(DEV),
            sd["encoder_norm.bias"].to(DEV))
    outs = {}
    prev = ops.TS_FUSED_LAYER
    try:
        for fused in (True, False):
            ops.TS_FUSED_LAYER = fused
            h, img = ops.ts_encoder_forward_bf16(*args, drop_p=drop, seed=21, want_seq_image=True)
            outs[fused] = (h.clone(), img.clone())
    finally:
        ops.TS_FUSED_LAYER = prev
    assert torch.isfinite(outs[True][0]).all()
    assert torch.equal(outs[True][0], outs[False][0])
    # the image's padding rows (nodes .. 128-row boundary) are never written: compare what the Gram GEMM reads from it
    assert torch.equal(ops.tc_cosine_gram(outs[True][1], B, N, P), ops.tc_cosine_gram(outs[False][1], B, N, P))


@pytest.mark.parametrize("B,N,P,drop", [(2, 9, 168, 0.0), (1, 11, 336, 0.0), (3, 5, 24, 0.0), (2, 13, 168, 0.1), (4, 40, 168, 0.1)])
def test_fused_ffn_kernel_matches_two_launch_path(B, N, P, drop):
    """tc_ffn_kernel (linear1 + ReLU + dropout + linear2 + dropout + residual + LayerNorm in one launch, the hidden
    activations never leave the SM; opt-in with STEP_B200_FFN_FUSED=1) vs the default tc_linear_kernel<RELU_IMG> +
    tc_linear_kernel<RESLN>: same MMAs in the same K order and the same dropout counters -> bit-identical hidden states
    and Gram operand image."""
    import os
    from step_b200 import ops
    sd = O.synthetic_tsformer_params(2)
    g = torch.Generator().manual_seed(9)
    series = torch.randn(B, P * 12, N, generator=g).to(DEV)
    layers = _layers(sd)
    images = ops.ts_pack_layer_images(layers)
    args = (series, sd["patch_embedding.input_embedding.weight"].to(DEV), sd["patch_embedding.input_embedding.bias"].to(DEV),
            sd["positional_encoding.position_embedding"].to(DEV), layers, images, sd["encoder_norm.weight"].to(DEV),
            sd["encoder_norm.bias"].to(DEV))
    outs = {}
    try:
        for fused in ("1", "0"):
            os.environ["STEP_B200_FFN_FUSED"] = fused
            h, img = ops.ts_encoder_forward_bf16(*args, drop_p=drop, seed=21, want_seq_image=True)
            outs[fused] = (h.clone(), img.clone())
    finally:
        os.environ.pop("STEP_B200_FFN_FUSED", None)
    assert torch.isfinite(outs["1"][0]).all()
    assert torch.equal(outs["1"][0], outs["0"][0])
    assert torch.equal(ops.tc_cosine_gram(outs["1"][1], B, N, P), ops.tc_cosine_gram(outs["0"][1], B, N, P))
