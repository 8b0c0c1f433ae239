"""Dropout-live distribution checks, one per stochastic site of the timed configuration (the reference leaves the frozen
TSFormer in train(), so its five dropout sites per layer and the positional dropout are live during STEP training; the gcn
dropout of Graph WaveNet is live too).  Each test isolates one site through the C ABI and checks: keep probability 1 - p
(binomial tolerance), kept values scaled by 1/(1-p), reproducible per seed, different across seeds.  The two attention-
probability sites are in test_gpu_kernels.py / test_gpu_tc.py."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
P_DROP = 0.1


def _check_mask(dropped, reference, p, what):
    """dropped / reference: same-shape tensors; entries where reference != 0 must be either 0 or reference / (1 - p)."""
    live = reference.abs() > 1e-6
    n = int(live.sum())
    kept = live & (dropped != 0)
    frac = float(kept.sum()) / n
    tol = 5 * math.sqrt(p * (1 - p) / n) + 1e-4
    assert abs(frac - (1 - p)) < tol, (what, frac, n)
    ratio = dropped[kept] / reference[kept]
    assert (ratio - 1 / (1 - p)).abs().max().item() < 2e-2, what           # bf16 outputs: 2^-8 relative rounding
    return frac


def test_positional_dropout_fp32_and_bf16_embed():
    from step_b200 import ops
    g = torch.Generator().manual_seed(0)
    B, N, P = 2, 37, 168
    series = torch.randn(B, P * 12, N, generator=g).to(DEV)
    w, b, pos = torch.randn(96, 12, generator=g).to(DEV) * 0.1, torch.randn(96, generator=g).to(DEV) * 0.1, torch.randn(P, 96, generator=g).to(DEV) * 0.03
    ref = ops.ts_embed(series, w, b, pos)
    a = ops.ts_embed(series, w, b, pos, drop_p=P_DROP, seed=3)
    _check_mask(a.cpu(), ref.cpu(), P_DROP, "ts_embed fp32")
    assert torch.equal(a, ops.ts_embed(series, w, b, pos, drop_p=P_DROP, seed=3))
    assert not torch.equal(a, ops.ts_embed(series, w, b, pos, drop_p=P_DROP, seed=4))
    # bf16 path: tc_embed_kernel into the tile image
    L = ops._L()
    T = B * N * P
    st = ops._enter(series)
    imgs = []
    for p_, seed in ((0.0, 3), (P_DROP, 3), (P_DROP, 3), (P_DROP, 4)):
        img = torch.zeros(((T + 127) // 128) * 96 * 256, device=DEV, dtype=torch.uint8)
        sB, sT, sN = series.stride()
        ops.check(L.step_tc_embed_fwd(series.data_ptr(), sB, sT, sN, B, N, P, w.data_ptr(), b.data_ptr(), pos.data_ptr(),
                                      img.data_ptr(), p_, seed, st), "step_tc_embed_fwd")
        imgs.append(ops.tc_image_to_rows(img, T, 96).cpu())
    _check_mask(imgs[1], imgs[0], P_DROP, "tc_embed bf16")
    assert torch.equal(imgs[1], imgs[2]) and not torch.equal(imgs[1], imgs[3])


def test_ffn_hidden_dropout_fp32_and_bf16():
    from step_b200 import ops
    g = torch.Generator().manual_seed(1)
    T = 4000
    x, w, b = torch.randn(T, 96, generator=g).to(DEV), (torch.randn(384, 96, generator=g) * 0.1).to(DEV), torch.randn(384, generator=g).to(DEV) * 0.1
    ref = ops.linear(x, w, b, epilogue=1)
    a = ops.linear(x, w, b, epilogue=1, drop_p=P_DROP, seed=5)
    _check_mask(a.cpu(), ref.cpu(), P_DROP, "linear relu fp32")
    assert torch.equal(a, ops.linear(x, w, b, epilogue=1, drop_p=P_DROP, seed=5))
    assert not torch.equal(a, ops.linear(x, w, b, epilogue=1, drop_p=P_DROP, seed=6))
    x_img, w_img = ops.tc_rows_to_image(x), ops.tc_pack_weight(w)
    rows = lambda img: ops.tc_image_to_rows(img, T, 384).cpu()
    ref16 = rows(ops.tc_linear(x_img, w_img, b, T, 96, 384, 1)[0])
    a16 = rows(ops.tc_linear(x_img, w_img, b, T, 96, 384, 1, drop_p=P_DROP, seed=5)[0])
    _check_mask(a16, ref16, P_DROP, "tc_linear relu bf16")
    assert torch.equal(a16, rows(ops.tc_linear(x_img, w_img, b, T, 96, 384, 1, drop_p=P_DROP, seed=5)[0]))
    assert not torch.equal(a16, rows(ops.tc_linear(x_img, w_img, b, T, 96, 384, 1, drop_p=P_DROP, seed=6)[0]))


@pytest.mark.parametrize("K", [96, 384])          # out-projection (dropout1) and FFN2 (dropout2)
def test_residual_branch_dropout_fp32_and_bf16(K):
    """LayerNorm(residual + drop(C)) hides the mask behind a normalisation, so C is made constant (A = 1, W = 1/K, zero
    residual): every row of drop(C) then holds the two values {0, 1/(1-p)} and the dropped entries are exactly the ones
    below the row mean after LayerNorm."""
    from step_b200 import ops
    T = 4096
    a = torch.ones(T, K, device=DEV)
    w = torch.full((96, K), 1.0 / K, device=DEV)
    zero, one = torch.zeros(96, device=DEV), torch.ones(96, device=DEV)
    res = torch.zeros(T, 96, device=DEV)
    y = ops.linear(a, w, zero, epilogue=2, residual=res, ln_w=one, ln_b=zero, drop_p=P_DROP, seed=8).cpu()
    dropped = (y < 0).float().mean().item()
    assert abs(dropped - P_DROP) < 5 * math.sqrt(P_DROP * (1 - P_DROP) / y.numel()) + 1e-4
    assert torch.equal(y, ops.linear(a, w, zero, epilogue=2, residual=res, ln_w=one, ln_b=zero, drop_p=P_DROP, seed=8).cpu())
    a_img, w_img, r_img = ops.tc_rows_to_image(a), ops.tc_pack_weight(w), ops.tc_rows_to_image(res)
    rows = lambda out: ops.tc_image_to_rows(out[0], T, 96).cpu()
    y16 = rows(ops.tc_linear(a_img, w_img, zero, T, K, 96, 2, res_img=r_img, ln_w=one, ln_b=zero, drop_p=P_DROP, seed=8))
    d16 = (y16 < 0).float().mean().item()
    assert abs(d16 - P_DROP) < 5 * math.sqrt(P_DROP * (1 - P_DROP) / y16.numel()) + 1e-4
    y16b = rows(ops.tc_linear(a_img, w_img, zero, T, K, 96, 2, res_img=r_img, ln_w=one, ln_b=zero, drop_p=P_DROP, seed=9))
    assert not torch.equal(y16, y16b)
    # per-row counts follow a binomial(96, p): variance check guards against correlated draws along a row
    cnt = (y16 < 0).float().sum(1)
    assert cnt.var().item() == pytest.approx(96 * P_DROP * (1 - P_DROP), rel=0.15)


def test_gwnet_gcn_dropout_site():
    """The gcn-output dropout of the Graph WaveNet layers (p = 0.3, Philox): the stand-alone probe applies exactly the layer
    kernels' mask function / key derivation."""
    from step_b200 import ops
    L = ops._L()
    rows, p = 20000, 0.3
    x = torch.ones(rows, 32, device=DEV)
    st = ops._enter(x)
    outs = {}
    for seed, layer in ((1, 0), (1, 0), (1, 1), (2, 0)):
        y = torch.empty_like(x)
        ops.check(L.step_gwnet_dropout_probe(x.data_ptr(), rows, p, seed, layer, y.data_ptr(), st), "probe")
        outs.setdefault((seed, layer), []).append(y.cpu())
    a = outs[(1, 0)][0]
    _check_mask(a, torch.ones(rows, 32), p, "gwnet gcn dropout")
    assert torch.equal(a, outs[(1, 0)][1])
    assert not torch.equal(a, outs[(1, 1)][0]) and not torch.equal(a, outs[(2, 0)][0])      # per-layer and per-seed streams
    # independence across layers: the joint keep rate of two layers' masks is (1-p)^2
    both = ((a != 0) & (outs[(1, 1)][0] != 0)).float().mean().item()
    assert abs(both - (1 - p) ** 2) < 5e-3
