"""GPU: tcgen05 GEMM / implicit-GEMM conv against a plain PyTorch fp32 reference of the same op
(inputs rounded to fp16 first, so the only difference is fp32 accumulation order and the fp16
rounding of the output).  Tolerance: |err| <= 2e-2 * max|ref| for fp16 outputs, 2e-3 for fp32."""
import pytest
import torch
import torch.nn.functional as F

from dwg import ops

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _rel(got, ref):
    return float((got.float() - ref).abs().max() / (ref.abs().max() + 1e-12))


@pytest.mark.parametrize('M,N,K', [(128, 128, 64), (256, 128, 128), (8192, 320, 320), (300, 77, 40), (4096, 1280, 2560),
                                   (130, 250, 72), (2, 1280, 320)])
def test_gemm_plain(M, N, K):
    torch.manual_seed(0)
    a = torch.randn(M, K, device=DEV).half()
    b = torch.randn(N, K, device=DEV).half()
    ref = a.float() @ b.float().t()
    c = ops.gemm(a, b, out_dtype=torch.float32)
    assert _rel(c, ref) < 2e-3
    bias = torch.randn(N, device=DEV)
    res = torch.randn(M, N, device=DEV).half()
    c2 = ops.gemm(a, b, bias=bias, residual=res, alpha=0.5, act='silu')
    ref2 = F.silu(0.5 * ref + bias) + res.float()
    assert c2.dtype == torch.float16 and _rel(c2, ref2) < 2e-2
    c3 = ops.gemm(a, b, bias=bias, act='gelu', out_dtype=torch.float32)
    assert _rel(c3, F.gelu(ref + bias)) < 2e-3


def test_gemm_strided_batched_attention_shapes():
    """QK^T and PV for 8 heads x 2 samples with head dim 40, straight from [B, T, heads*hd] tensors."""
    torch.manual_seed(1)
    B, T, Hh, hd, Tk = 2, 1024, 8, 40, 77
    q = torch.randn(B, T, Hh * hd, device=DEV).half()
    k = torch.randn(B, Tk, Hh * hd, device=DEV).half()
    q4 = q.view(B, T, Hh, hd).permute(0, 2, 1, 3)          # [B, heads, T, hd] strided view
    k4 = k.view(B, Tk, Hh, hd).permute(0, 2, 1, 3)
    s = ops.gemm(q4, k4, alpha=hd ** -0.5, out_dtype=torch.float32)
    ref = torch.einsum('bhtd,bhsd->bhts', q4.float(), k4.float()) * hd ** -0.5
    assert s.shape == (B, Hh, T, Tk) and _rel(s, ref) < 2e-3
    # PV with V^T [B, heads, hd, Tk] (K = Tk = 80 after padding to a multiple of 8)
    p = torch.softmax(ref, -1)
    Tkp = 80
    pp = torch.zeros(B, Hh, T, Tkp, device=DEV, dtype=torch.float16)
    pp[..., :Tk] = p.half()
    vt = torch.zeros(B, Hh, hd, Tkp, device=DEV, dtype=torch.float16)
    vt[..., :Tk] = torch.randn(B, Hh, hd, Tk, device=DEV).half()
    out = torch.empty(B, T, Hh * hd, device=DEV, dtype=torch.float16)
    o4 = out.view(B, T, Hh, hd).permute(0, 2, 1, 3)
    ops.gemm(pp, vt, out=o4)
    ref_o = torch.einsum('bhts,bhds->bhtd', pp.float(), vt.float())
    assert _rel(o4, ref_o) < 2e-2


@pytest.mark.parametrize('N,H,W,Cin,Cout,k,stride', [
    (2, 64, 64, 320, 320, 3, 1), (2, 32, 32, 640, 640, 3, 1), (2, 16, 16, 1280, 1280, 3, 1), (2, 8, 8, 1280, 1280, 3, 1),
    (2, 64, 64, 320, 320, 3, 2), (2, 64, 64, 8, 320, 3, 1), (2, 64, 64, 320, 8, 3, 1), (2, 32, 32, 640, 320, 1, 1),
    (1, 128, 128, 128, 128, 3, 1), (1, 256, 256, 16, 32, 3, 2), (3, 8, 8, 64, 96, 3, 1), (1, 24, 40, 64, 64, 3, 1)])
def test_conv2d_nhwc_vs_torch(N, H, W, Cin, Cout, k, stride):
    torch.manual_seed(2)
    x = torch.randn(N, H, W, Cin, device=DEV).half()
    w = (torch.randn(Cout, k, k, Cin, device=DEV) / (k * k * Cin) ** 0.5).half()
    bias = torch.randn(Cout, device=DEV)
    pad = k // 2
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), bias, stride=stride, padding=pad).permute(0, 2, 3, 1)
    y = ops.conv2d_nhwc(x, w, bias=bias, stride=stride, padding=pad, out_dtype=torch.float32)
    assert y.shape == ref.shape and _rel(y, ref) < 2e-3
    temb = torch.randn(N, Cout, device=DEV)
    res = torch.randn_like(ref).half()
    y2 = ops.conv2d_nhwc(x, w, bias=bias, bias2=temb, residual=res, stride=stride, padding=pad)
    ref2 = ref + temb[:, None, None, :] + res.float()
    assert _rel(y2, ref2) < 2e-2


def test_conv2d_asymmetric_padding_downsample():
    """VAE encoder downsample: F.pad(x, (0,1,0,1)) then 3x3 stride-2 conv with padding 0."""
    torch.manual_seed(3)
    N, H, W, C = 1, 64, 64, 128
    x = torch.randn(N, H, W, C, device=DEV).half()
    w = (torch.randn(C, 3, 3, C, device=DEV) / (9 * C) ** 0.5).half()
    xp = F.pad(x.float().permute(0, 3, 1, 2), (0, 1, 0, 1))
    ref = F.conv2d(xp, w.float().permute(0, 3, 1, 2), None, stride=2, padding=0).permute(0, 2, 3, 1)
    y = ops.conv2d_nhwc(x, w, stride=2, padding=(0, 0), out_hw=(H // 2, W // 2), out_dtype=torch.float32)
    assert _rel(y, ref) < 2e-3


@pytest.mark.parametrize('B,heads,T,Tk,hd', [(2, 8, 4096, 4096, 40), (2, 8, 1024, 1024, 80), (2, 8, 4096, 77, 40), (2, 8, 1024, 77, 80),
                                            (1, 4, 256, 200, 64), (2, 8, 256, 256, 8), (1, 2, 100, 300, 128), (2, 8, 64, 77, 32)])
def test_fused_attention_vs_sdpa(B, heads, T, Tk, hd):
    torch.manual_seed(5)
    C = heads * hd
    q = torch.randn(B, T, C, device=DEV).half()
    k = torch.randn(B, Tk, C, device=DEV).half()
    v = torch.randn(B, Tk, C, device=DEV).half()
    Tkp = (Tk + 7) // 8 * 8
    vt = torch.zeros(B, C, Tkp, device=DEV, dtype=torch.float16)
    vt[:, :, :Tk] = v.transpose(1, 2)
    out = ops.attention(q, k, vt, heads, Tk)
    sp = lambda t: t.float().view(B, -1, heads, hd).transpose(1, 2)
    ref = F.scaled_dot_product_attention(sp(q), sp(k), sp(v)).transpose(1, 2).reshape(B, T, C)
    assert _rel(out, ref) < 2e-2


def test_gemm_fused_geglu_epilogue():
    torch.manual_seed(7)
    M, K, inner = 2048, 640, 2560
    a = torch.randn(M, K, device=DEV).half()
    w = (torch.randn(2 * inner, K, device=DEV) / K ** 0.5).half()
    bias = torch.randn(2 * inner, device=DEV) * 0.1
    full = a.float() @ w.float().t() + bias
    ref = full[:, :inner] * F.gelu(full[:, inner:])
    wi = torch.stack([w[:inner], w[inner:]], dim=1).reshape(2 * inner, K).contiguous()
    bi = torch.stack([bias[:inner], bias[inner:]], dim=1).reshape(2 * inner).contiguous()
    out = ops.gemm(a, wi, bias=bi, act='geglu')
    assert out.shape == (M, inner) and _rel(out, ref) < 2e-2


def test_gemm_tiny_m_and_deep_split_k():
    torch.manual_seed(8)
    for (M, N, K) in ((2, 1280, 320), (128, 1280, 5120), (128, 1280, 1280), (512, 1280, 1280)):
        a = torch.randn(M, K, device=DEV).half()
        b = (torch.randn(N, K, device=DEV) / K ** 0.5).half()
        res = torch.randn(M, N, device=DEV).half()
        c = ops.gemm(a, b, residual=res, out_dtype=torch.float32)
        assert _rel(c, a.float() @ b.float().t() + res.float()) < 2e-3


@pytest.fixture
def force_pair():
    """Force the CTA-pair (tcgen05.mma.cta_group::2) path wherever it is legal, restore the planner afterwards."""
    from dwg._lib import lib
    lib().dwg_gemm_tune_pair(1)
    yield lib()
    lib().dwg_gemm_tune_pair(-1)
    lib().dwg_gemm_tune(0, 0)


@pytest.mark.parametrize('M,N,K,bn,ks', [(256, 128, 64, 0, 0), (8192, 320, 320, 0, 0), (4096, 1280, 2560, 256, 1), (512, 1280, 1280, 128, 1),
                                         (2048, 640, 640, 160, 1), (1024, 250, 72, 64, 1), (512, 1280, 5120, 128, 4), (8192, 4096, 512, 256, 1)])
def test_gemm_cta_pair_matches_torch(force_pair, M, N, K, bn, ks):
    L = force_pair
    L.dwg_gemm_tune(bn, ks)
    torch.manual_seed(0)
    a = torch.randn(M, K, device=DEV).half()
    b = torch.randn(N, K, device=DEV).half()
    ref = a.float() @ b.float().t()
    c = ops.gemm(a, b, out_dtype=torch.float32)
    assert L.dwg_gemm_last_pair() == 1, 'pair mode was not taken'
    assert _rel(c, ref) < 2e-3
    bias = torch.randn(N, device=DEV)
    res = torch.randn(M, N, device=DEV).half()
    c2 = ops.gemm(a, b, bias=bias, residual=res, alpha=0.5, act='silu')
    assert c2.dtype == torch.float16 and _rel(c2, F.silu(0.5 * ref + bias) + res.float()) < 2e-2
    # repeated launches (persistent tiles, barrier phases, TMEM stages) stay correct
    for _ in range(3):
        c = ops.gemm(a, b, out_dtype=torch.float32)
    assert _rel(c, ref) < 2e-3


@pytest.mark.parametrize('N,H,W,Cin,Cout,k,stride', [
    (2, 64, 64, 320, 320, 3, 1), (2, 32, 32, 640, 640, 3, 1), (2, 16, 16, 1280, 1280, 3, 1), (2, 64, 64, 320, 320, 3, 2),
    (1, 128, 128, 128, 128, 3, 1), (2, 32, 32, 640, 320, 1, 1), (1, 256, 256, 128, 256, 3, 1)])
def test_conv2d_cta_pair_matches_torch(force_pair, N, H, W, Cin, Cout, k, stride):
    L = force_pair
    torch.manual_seed(2)
    x = torch.randn(N, H, W, Cin, device=DEV).half()
    w = (torch.randn(Cout, k, k, Cin, device=DEV) / (k * k * Cin) ** 0.5).half()
    bias = torch.randn(Cout, device=DEV)
    pad = k // 2
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), bias, stride=stride, padding=pad).permute(0, 2, 3, 1)
    y = ops.conv2d_nhwc(x, w, bias=bias, stride=stride, padding=pad, out_dtype=torch.float32)
    assert L.dwg_gemm_last_pair() == 1, 'pair mode was not taken'
    assert y.shape == ref.shape and _rel(y, ref) < 2e-3
    temb = torch.randn(N, Cout, device=DEV)
    res = torch.randn_like(ref).half()
    y2 = ops.conv2d_nhwc(x, w, bias=bias, bias2=temb, residual=res, stride=stride, padding=pad)
    assert _rel(y2, ref + temb[:, None, None, :] + res.float()) < 2e-2


def test_geglu_cta_pair(force_pair):
    torch.manual_seed(3)
    M, K, inner = 2048, 640, 2560
    a = torch.randn(M, K, device=DEV).half()
    w = (torch.randn(2 * inner, K, device=DEV) / K ** 0.5).half()        # rows interleaved (value_i, gate_i)
    bias = torch.randn(2 * inner, device=DEV)
    y = ops.gemm(a, w, bias=bias, act='geglu')
    assert force_pair.dwg_gemm_last_pair() == 1
    full = a.float() @ w.float().t() + bias
    ref = full[:, 0::2] * F.gelu(full[:, 1::2])
    assert y.shape == (M, inner) and _rel(y, ref) < 2e-2


@pytest.mark.parametrize('pair', [0, 1])
@pytest.mark.parametrize('N,H,W,Cin,Cout', [(1, 32, 32, 64, 64), (2, 64, 64, 320, 320), (1, 128, 128, 128, 128), (1, 64, 64, 8, 32),
                                            (2, 16, 16, 1280, 640), (1, 48, 24, 72, 96), (3, 16, 8, 64, 320)])
def test_conv3x3_halo_mode_matches_torch(pair, N, H, W, Cin, Cout):
    """Halo mode: one (16+2)x(8+2) activation halo per 64-channel slice feeds the nine taps through shifted descriptors."""
    from dwg._lib import lib
    L = lib()
    if pair and ((W // 8) * (H // 16) * N) % 2:
        pytest.skip('CTA pairs need an even number of 8x16-pixel tiles (the planner falls back to the per-tap path)')
    torch.manual_seed(4)
    x = torch.randn(N, H, W, Cin, device=DEV).half()
    w = (torch.randn(Cout, 3, 3, Cin, device=DEV) / (9 * Cin) ** 0.5).half()
    bias = torch.randn(Cout, device=DEV)
    temb = torch.randn(N, Cout, device=DEV)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), bias, padding=1).permute(0, 2, 3, 1)
    res = torch.randn_like(ref).half()
    L.dwg_gemm_tune_halo(1, 0)
    L.dwg_gemm_tune_pair(pair)
    L.dwg_gemm_tune(128, 1)                          # halo mode does not split K
    try:
        y = ops.conv2d_nhwc(x, w, bias=bias, out_dtype=torch.float32)
        took = L.dwg_gemm_last_halo()
        y2 = ops.conv2d_nhwc(x, w, bias=bias, bias2=temb, residual=res)
        for _ in range(2):
            y3 = ops.conv2d_nhwc(x, w, bias=bias, out_dtype=torch.float32)
    finally:
        L.dwg_gemm_tune_halo(-1, 0)
        L.dwg_gemm_tune_pair(-1)
        L.dwg_gemm_tune(0, 0)
    assert took == 1, 'halo mode was not taken'
    assert _rel(y, ref) < 2e-3 and _rel(y3, ref) < 2e-3
    assert _rel(y2, ref + temb[:, None, None, :] + res.float()) < 2e-2


@pytest.mark.parametrize('case', ['gemm', 'gemm_res', 'conv3', 'conv1_res', 'conv_small', 'conv_s2'])
def test_epilogue_column_statistics_feed_groupnorm(case):
    """GroupNorm statistics from the PRODUCING epilogue: the per-column fixed-point sums a GEMM / conv launch accumulates equal
    the sums of the fp16 tensor it wrote (exactly: integer accumulation of the rounded values), and dwg_groupnorm_apply_cs on
    them equals the two-pass GroupNorm bit for bit."""
    import torch
    from dwg import ops
    torch.manual_seed(7)
    dev = 'cuda'
    ops.STATS_ARENA.reset(dev)
    if case.startswith('gemm'):
        M, N, K, rows = 2 * 1024, 320, 192, 1024
        a, b = torch.randn(M, K, device=dev).half(), (torch.randn(N, K, device=dev) * 0.1).half()
        res = torch.randn(M, N, device=dev).half() if case == 'gemm_res' else None
        y = ops.gemm(a, b, bias=torch.randn(N, device=dev), residual=res, colstats_rows=rows)
        groups, C = M // rows, N
        yv = y.view(groups, rows, C)
    else:
        Nimg, H, W, Cin, Cout, k, stride = {'conv3': (2, 32, 32, 64, 128, 3, 1), 'conv1_res': (2, 16, 16, 320, 320, 1, 1),
                                            'conv_small': (2, 8, 8, 128, 256, 3, 1), 'conv_s2': (1, 64, 64, 32, 64, 3, 2)}[case]
        x = torch.randn(Nimg, H, W, Cin, device=dev).half()
        w = (torch.randn(Cout, k, k, Cin, device=dev) * 0.05).half()
        Ho = H // stride
        res = torch.randn(Nimg, Ho, Ho, Cout, device=dev).half() if case == 'conv1_res' else None
        y = ops.conv2d_nhwc(x, w, bias=torch.randn(Cout, device=dev), residual=res, stride=stride, padding=1 if k == 3 else 0, stats=True)
        groups, C = Nimg, Cout
        yv = y.view(groups, -1, C)
    cs = y._cs
    assert cs.shape == (4, groups, C, 2) and cs.dtype == torch.int64
    yd = yv.double()
    ref = torch.stack([yd.sum(1), (yd * yd).sum(1)], dim=-1)                       # [groups, C, 2]
    got = cs.sum(0).double() / 2.0 ** 20
    # each 32-row partial is converted to 2^-20 fixed point once: |error| <= 2^-21 per partial + fp32 rounding of the 32-term partial sums
    torch.testing.assert_close(got, ref, rtol=2e-6, atol=1e-3)
    g, bta = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev) * 0.1
    y4 = y.view(groups, -1, 1, C) if y.dim() == 2 else y
    two_pass, st_ref = ops.group_norm(y4.contiguous(), g, bta, 32, 1e-5, silu=True, return_stats=True)
    one_pass, st = ops.group_norm(y4.contiguous(), g, bta, 32, 1e-5, silu=True, return_stats=True, colstats=cs)
    assert (one_pass.float() - two_pass.float()).abs().max() <= 2e-3 * two_pass.float().abs().max()      # statistics agree to ~1e-7: a few fp16 ulps at most
    torch.testing.assert_close(st.double() / 2.0 ** 20, st_ref.double() / 2.0 ** 20, rtol=2e-6, atol=1e-2)


def test_gemm_broadcast_a_over_batch():
    """a_b1 == 0: one A matrix (a weight) for every batch entry -- V^T[b] = Wv ctx[b]^T of the attention layers in ONE launch."""
    torch.manual_seed(3)
    B, M, N, K = 2, 320, 1024, 320
    w = (torch.randn(M, K, device=DEV) * 0.1).half()
    x = torch.randn(B, N, K, device=DEV).half()
    out = torch.zeros(B, M, N + 8, device=DEV, dtype=torch.float16)
    ops.gemm(w.unsqueeze(0).expand(B, -1, -1), x, out=out[:, :, :N])
    ref = torch.einsum('mk,bnk->bmn', w.float(), x.float())
    err = (out[:, :, :N].float() - ref).abs().max() / ref.abs().max()
    assert float(err) < 2e-3, float(err)
    assert float(out[:, :, N:].abs().max()) == 0.0


def _ln_ref(x, g, b, eps=1e-5):
    return F.layer_norm(x.float(), (x.shape[-1],), g, b, eps)


@pytest.mark.parametrize('M,C,N', [(512, 320, 640), (8192, 320, 640), (300, 1280, 2560)])
def test_gemm_layernorm_folded_rows_and_rowstats(M, C, N):
    """LayerNorm folded into the consuming GEMM (ln_mode 1) from the row statistics the PRODUCING GEMM left in its epilogue,
    against torch: y = LN(x) W^T + bias with x = fp16(A0 W0^T + r).  Also the GEGLU epilogue behind the folded LN."""
    torch.manual_seed(M + C)
    a0 = torch.randn(M, C, device=DEV).half()
    w0 = (torch.randn(C, C, device=DEV) / C ** 0.5).half()
    r = (torch.randn(M, C, device=DEV) + 0.7).half()                          # a mean offset: the cancellation the fold must survive
    x = ops.gemm_ln(a0, w0, residual=r, rowstats=True)
    xr = (a0.float() @ w0.float().t() + r.float()).half()
    assert (x.float() - xr.float()).abs().max() < 2e-2
    rs = x._rs.double() / 2 ** 20
    assert torch.allclose(rs[:, 0], x.double().sum(1), atol=1e-3) and torch.allclose(rs[:, 1], (x.double() ** 2).sum(1), rtol=1e-5, atol=1e-3)
    g, b = torch.rand(C, device=DEV) + 0.5, torch.randn(C, device=DEV) * 0.2
    W = torch.randn(N, C, device=DEV) / C ** 0.5
    bias = torch.randn(N, device=DEV) * 0.1
    wp = (W * g[None]).half()
    c1, c2 = wp.float().sum(1), W @ b + bias
    y = ops.gemm_ln(x, wp, bias=c2, ln=(x._rs, c1, C, 1e-5))
    ref = _ln_ref(x, g, b) @ W.t() + bias
    err = (y.float() - ref).abs().max() / ref.abs().max()
    assert float(err) < 4e-3, float(err)
    # GEGLU: interleaved (value, gate) rows
    inner = N // 2
    Wi = torch.stack([W[:inner], W[inner:]], 1).reshape(N, C)
    bi = torch.stack([bias[:inner], bias[inner:]], 1).reshape(N)
    wpi = (Wi * g[None]).half()
    yg = ops.gemm_ln(x, wpi, bias=Wi @ b + bi, act='geglu', ln=(x._rs, wpi.float().sum(1), C, 1e-5))
    refg = ref[:, :inner] * F.gelu(ref[:, inner:])
    errg = (yg.float() - refg).abs().max() / refg.abs().max()
    assert float(errg) < 6e-3, float(errg)


def test_gemm_layernorm_folded_columns_batched():
    """ln_mode 2: V^T[b] = Wv LN(x[b])^T with the tokens as the B rows (per-column statistics), A broadcast over the batch."""
    torch.manual_seed(9)
    B, T, C = 2, 1024, 320
    x = (torch.randn(B * T, C, device=DEV) * 1.3 + 0.4).half()
    ident = torch.eye(C, device=DEV).half()
    xs = ops.gemm_ln(x, ident, rowstats=True)                                  # copy through the epilogue: row statistics of x
    assert torch.equal(xs, x)
    g, b = torch.rand(C, device=DEV) + 0.5, torch.randn(C, device=DEV) * 0.2
    Wv = torch.randn(C, C, device=DEV) / C ** 0.5
    wp = (Wv * g[None]).half()
    vT = torch.empty(B, C, T, device=DEV, dtype=torch.float16)
    ops.gemm_ln(wp.unsqueeze(0).expand(B, -1, -1), x.view(B, T, C), ln_cols=(xs._rs, wp.float().sum(1), Wv @ b, C, 1e-5), out=vT)
    ref = torch.einsum('ck,btk->bct', Wv, _ln_ref(x, g, b).view(B, T, C))
    err = (vT.float() - ref).abs().max() / ref.abs().max()
    assert float(err) < 4e-3, float(err)
