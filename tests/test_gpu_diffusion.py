"""GPU: the diffusion blocks (UNet, ControlNet, VAE encoder fwd + input-gradient bwd, SDS algebra)
on dwg kernels against the fp32 CPU oracle, on a reduced-width model with the full topology
(seconds; tests/test_gpu_diffusion_sd15.py repeats the comparison at the benchmark's SD1.5 sizes and guidance
scale 50).  fp16 tensor-core arithmetic (fp32 accumulate) vs an fp32 oracle: stated tolerance rel-L2 <= 5e-3 on
eps / latents (measured ~2e-3), <= 1e-2 on the VAE input gradient."""
import numpy as np
import pytest
import torch

from dwg import ops
from dwg.diffusion import guidance as G, model as M, weights as W

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def rel_l2(got, ref):
    return float((got.float().cpu() - ref).norm() / (ref.norm() + 1e-12))


def test_norm_kernels_vs_torch():
    torch.manual_seed(0)
    x = torch.randn(2, 16, 16, 320, device=DEV).half()
    g, b = torch.rand(320, device=DEV) + 0.5, torch.randn(320, device=DEV) * 0.1
    ref = torch.nn.functional.group_norm(x.float().permute(0, 3, 1, 2), 32, g, b, 1e-5)
    y = ops.group_norm(x, g, b, 32, 1e-5, silu=False)
    assert rel_l2(y.permute(0, 3, 1, 2), ref.cpu()) < 1e-2, rel_l2(y.permute(0, 3, 1, 2), ref.cpu())
    y2 = ops.group_norm(x, g, b, 32, 1e-5, silu=True)
    e2 = rel_l2(y2.permute(0, 3, 1, 2), torch.nn.functional.silu(ref).cpu())
    assert e2 < 1e-2, e2
    # backward
    xr = x.float().requires_grad_(True)
    out = torch.nn.functional.silu(torch.nn.functional.group_norm(xr.permute(0, 3, 1, 2), 32, g, b, 1e-5)).permute(0, 2, 3, 1)
    dy = torch.randn(out.shape, device=DEV).half()
    out.backward(dy.float())
    _, st = ops.group_norm(x, g, b, 32, 1e-5, silu=True, return_stats=True)
    dx = ops.group_norm_bwd(x, dy, st, g, b, 32, 1e-5, True)
    e3 = rel_l2(dx, xr.grad.cpu())
    assert e3 < 2e-2, ('gn bwd', e3)
    # layernorm / softmax / geglu
    t = torch.randn(50, 640, device=DEV).half()
    e4 = rel_l2(ops.layer_norm(t, g.repeat(2), b.repeat(2)), torch.nn.functional.layer_norm(t.float(), (640,), g.repeat(2), b.repeat(2)).cpu())
    assert e4 < 1e-2, ('ln', e4)
    for cols, pad in ((77, 80), (4096, 4096), (300, 304)):
        s = torch.randn(33, pad, device=DEV).half()
        ref = torch.softmax(s.float()[:, :cols], -1)
        ops.softmax_rows_(s, cols)
        e5 = rel_l2(s[:, :cols], ref.cpu())
        assert e5 < 1e-2 and float(s[:, cols:].float().abs().sum()) == 0, ('softmax', cols, e5, float(s[:, cols:].float().abs().sum()))
    gg = torch.randn(64, 2560, device=DEV).half()
    a, bb = gg.float().chunk(2, -1)
    e6 = rel_l2(ops.geglu(gg), (a * torch.nn.functional.gelu(bb)).cpu())
    assert e6 < 1e-2, ('geglu', e6)


def _tiny():
    cfg = W.TINY
    return cfg, W.make_unet(cfg), W.make_controlnet(cfg), W.make_vae_encoder(W.TINY_VAE)


def test_unet_and_controlnet_match_oracle():
    from oracle import diffusion as od
    cfg, u_sd, c_sd, _ = _tiny()
    torch.manual_seed(1)
    x = torch.randn(2, 4, 16, 16)
    t = torch.tensor([487])
    ctx = torch.randn(2, 77, cfg['ctx_dim'])
    cond = torch.rand(2, 3, 128, 128)
    with torch.no_grad():
        down_r, mid_r = od.controlnet_forward(c_sd, cfg, x, t, ctx, cond)
        eps_r = od.unet_forward(u_sd, cfg, x, t, ctx, down_r, mid_r)
        eps_plain = od.unet_forward(u_sd, cfg, x, t, ctx)
    cn, un = M.ControlNet(c_sd, cfg, DEV), M.UNet(u_sd, cfg, DEV)
    down, mid = cn.forward(x.to(DEV), t.to(DEV), ctx.to(DEV), cond.to(DEV))
    for i, (d, r) in enumerate(zip(down, down_r)):
        assert rel_l2(d.permute(0, 3, 1, 2), r) < 5e-3, i
    assert rel_l2(mid.permute(0, 3, 1, 2), mid_r) < 5e-3
    eps = un.forward(x.to(DEV), t.to(DEV), ctx.to(DEV), down, mid)
    assert eps.shape == eps_r.shape and rel_l2(eps, eps_r) < 5e-3
    assert rel_l2(un.forward(x.to(DEV), t.to(DEV), ctx.to(DEV)), eps_plain) < 5e-3


def test_vae_encode_forward_and_input_gradient_match_oracle():
    from oracle import diffusion as od
    _, _, _, v_sd = _tiny()
    vcfg = W.TINY_VAE
    torch.manual_seed(2)
    img = torch.rand(1, 3, 64, 64)
    eps = torch.randn(1, 4, 8, 8)
    img_r = img.clone().requires_grad_(True)
    lat_r = od.vae_encode_latents(v_sd, vcfg, img_r, eps)
    gl = torch.randn_like(lat_r)
    (lat_r * gl).sum().backward()
    enc = M.VAEEncoder(v_sd, vcfg, DEV)
    img_g = img.to(DEV).requires_grad_(True)
    lat = M.vae_encode(enc, img_g, eps.to(DEV))
    assert rel_l2(lat, lat_r.detach()) < 5e-3
    (lat * gl.to(DEV)).sum().backward()
    assert rel_l2(img_g.grad, img_r.grad) < 1e-2


def test_sds_step_matches_oracle_and_specify_gradient():
    from oracle import diffusion as od
    cfg, u_sd, c_sd, v_sd = _tiny()
    torch.manual_seed(3)
    g = G.ControlNetScoreDistillation(u_sd, c_sd, v_sd, cfg, W.TINY_VAE, DEV, guidance_scale=7.5, default_image_size=128)
    img = torch.rand(1, 3, 128, 128)
    cond = torch.rand(1, 3, 128, 128)
    emb = {'neg': torch.randn(1, 77, cfg['ctx_dim']), 'text': torch.randn(1, 77, cfg['ctx_dim'])}
    t = torch.tensor([600])
    noise, veps = torch.randn(1, 4, 16, 16), torch.randn(1, 4, 16, 16)
    img_g = img.to(DEV).requires_grad_(True)
    out = g(img_g, {k: v.to(DEV) for k, v in emb.items()}, cond_inputs=cond.to(DEV), timestep=t.to(DEV), noise=noise.to(DEV),
            vae_eps=veps.to(DEV))
    # oracle
    img_r = img.clone().requires_grad_(True)
    lat_r = od.vae_encode_latents(v_sd, W.TINY_VAE, img_r, veps)
    with torch.no_grad():
        ln = od.add_noise(lat_r.detach(), noise, t)
        grad_r, np_r = od.sds_gradient(u_sd, c_sd, cfg, ln, noise, t, emb['neg'], emb['text'], cond, guidance_scale=7.5)
    assert rel_l2(out['latents'], lat_r.detach()) < 5e-3
    # CFG amplifies the rounding error of (eps_c - eps_u) by the guidance scale s: eps itself is held to 5e-3 (test above)
    assert rel_l2(out['noise_pred'], np_r) < 1.5e-2
    assert rel_l2(out['gradients'], grad_r) < 1.5e-2
    assert out['diffusion_loss'].shape == (1,) and float(out['diffusion_loss']) == 1.0
    out['diffusion_loss'].backward()
    (lat_r * grad_r).sum().backward()
    assert rel_l2(img_g.grad, img_r.grad) < 2e-2         # dominated by the CFG-amplified eps error above
    # CFG/SDS algebra alone (fp32 kernel): exact formula
    eu, ec, nz = torch.randn(3, 1, 4, 8, 8, device=DEV)
    gr, npd = ops.sds_grad(eu, ec, nz, 50.0, 1.0)
    torch.testing.assert_close(npd, eu + 50.0 * (ec - eu), rtol=1e-6, atol=1e-5)
    torch.testing.assert_close(gr, npd - nz, rtol=1e-6, atol=1e-5)


def test_prepare_head_start_and_two_streams_change_nothing():
    """guidance.prepare() (timestep / prompt / condition work enqueued early on the second stream) and the
    ControlNet-beside-UNet-encoder schedule are pure scheduling: same numbers as the single-stream path."""
    cfg, u_sd, c_sd, v_sd = _tiny()
    torch.manual_seed(5)
    img = torch.rand(1, 3, 128, 128, device=DEV)
    cond = torch.rand(1, 3, 128, 128, device=DEV)
    emb = {'neg': torch.randn(1, 77, cfg['ctx_dim'], device=DEV), 'text': torch.randn(1, 77, cfg['ctx_dim'], device=DEV)}
    t = torch.tensor([431], device=DEV)
    noise, veps = torch.randn(1, 4, 16, 16, device=DEV), torch.randn(1, 4, 16, 16, device=DEV)
    outs = []
    for mode in ('single', 'two_streams', 'prepared'):
        g = G.ControlNetScoreDistillation(u_sd, c_sd, v_sd, cfg, W.TINY_VAE, DEV, guidance_scale=7.5, default_image_size=128)
        g.two_streams = mode != 'single'
        x = img.clone().requires_grad_(True)
        if mode == 'prepared':
            g.prepare(emb, cond, timestep=t)
            out = g(x, emb, cond_inputs=cond, noise=noise, vae_eps=veps)          # timestep comes from prepare()
            assert int(out['timestep'][0]) == 431 and g._prepared is None
        else:
            out = g(x, emb, cond_inputs=cond, timestep=t, noise=noise, vae_eps=veps)
        out['diffusion_loss'].backward()
        torch.cuda.synchronize()
        outs.append((out['noise_pred'].clone(), x.grad.clone()))
    # Identical kernels and operands, and every reduction has a fixed order (GroupNorm statistics: integer atomics on
    # fixed-point partials; split-K: per-split slices added in split order), so the three schedules agree BITWISE --
    # a stream / PDL race would show up here.
    for npd, gr in outs[1:]:
        assert torch.equal(npd, outs[0][0]) and torch.equal(gr, outs[0][1])


@pytest.mark.parametrize('N,H,W,C', [(2, 64, 64, 320), (2, 8, 8, 1280), (2, 32, 32, 960), (2, 16, 16, 2560), (1, 64, 64, 512),
                                     (2, 8, 4, 128), (2, 128, 128, 128), (1, 512, 256, 128)])
def test_group_norm_one_launch_cluster_path_vs_torch(N, H, W, C):
    """UNet / ControlNet-sized tensors take the one-launch cluster kernel (DSMEM exchange of the partial sums); big tensors
    (last case) the two-pass kernels.  Both return the same raw (sum, sumsq) statistics the backward consumes."""
    from dwg._lib import lib
    torch.manual_seed(1)
    x = (torch.randn(N, H, W, C, device=DEV) * 1.7 + 0.3).half()
    g, b = torch.rand(C, device=DEV) + 0.5, torch.randn(C, device=DEV) * 0.1
    ref = torch.nn.functional.group_norm(x.float().permute(0, 3, 1, 2), 32, g, b, 1e-6)
    lib().dwg_groupnorm_set_fused(1)
    try:
        for silu in (False, True):
            y, st = ops.group_norm(x, g, b, 32, 1e-6, silu=silu, return_stats=True)
            fits = ((H * W + 7) // 8) * (4 * (C // 32)) * 2 + 32 * (C // 32) <= 200 * 1024
            assert lib().dwg_groupnorm_last_launches() == (1 if fits else 2)
            r = torch.nn.functional.silu(ref) if silu else ref
            assert rel_l2(y.permute(0, 3, 1, 2), r.cpu()) < 1e-2
    finally:
        lib().dwg_groupnorm_set_fused(0)
    xs = x.double().view(N, H * W, 32, C // 32)
    st_ref = torch.stack([xs.sum(dim=(1, 3)), (xs * xs).sum(dim=(1, 3))], dim=-1)          # [N, 32, 2]
    assert st.dtype == torch.int64                                                         # 2^-20 fixed point (include/dwg.h)
    torch.testing.assert_close(st.view(N, 32, 2).double().cpu() / 2.0 ** 20, st_ref.cpu(), rtol=2e-5, atol=1e-2)


def test_sd21_style_unet_controlnet_match_oracle():
    """cfg4 (SD2.1): per-level head counts (attention_head_dim = (5, 10, 20, 20) -> head width 64), OpenCLIP-width context,
    use_linear_projection (nn.Linear proj_in / proj_out) on a reduced-width model with the same topology."""
    from oracle import diffusion as od
    cfg = W.TINY21
    u_sd, c_sd = W.make_unet(cfg), W.make_controlnet(cfg)
    assert u_sd['down_blocks.0.attentions.0.proj_in.weight'].dim() == 2
    torch.manual_seed(11)
    x = torch.randn(2, 4, 16, 16)
    t = torch.tensor([333])
    ctx = torch.randn(2, 77, cfg['ctx_dim'])
    cond = torch.rand(2, 3, 128, 128)
    with torch.no_grad():
        down_r, mid_r = od.controlnet_forward(c_sd, cfg, x, t, ctx, cond)
        eps_r = od.unet_forward(u_sd, cfg, x, t, ctx, down_r, mid_r)
    cn, un = M.ControlNet(c_sd, cfg, DEV), M.UNet(u_sd, cfg, DEV)
    down, mid = cn.forward(x.to(DEV), t.to(DEV), ctx.to(DEV), cond.to(DEV))
    assert rel_l2(mid.permute(0, 3, 1, 2), mid_r) < 5e-3
    eps = un.forward(x.to(DEV), t.to(DEV), ctx.to(DEV), down, mid)
    assert rel_l2(eps, eps_r) < 5e-3, rel_l2(eps, eps_r)


def test_layernorm_folding_matches_layernorm_kernels(monkeypatch):
    """The folded transformer block (row statistics from the producing epilogue, normalisation applied algebraically in the
    consuming epilogue; off by default) against the LayerNorm-kernel path: same UNet, same inputs."""
    from dwg.diffusion import model as M, weights as Wt
    cfg = Wt.TINY
    sd = Wt.make_unet(cfg)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 4, 16, 16, generator=g).to(DEV)
    ctx = torch.randn(2, 77, cfg['ctx_dim'], generator=g).to(DEV)
    t = torch.tensor([321], device=DEV)
    outs = []
    for fold in (False, True):
        monkeypatch.setattr(M, 'LN_FOLD', fold)
        net = M.UNet(sd, cfg, DEV)
        assert bool(net._lnf) == fold
        outs.append(net.forward(x, t, ctx, None, None).float())
    rel = float((outs[1] - outs[0]).norm() / outs[0].norm())
    assert rel < 5e-3, rel


def test_group_norm_of_a_concatenation_without_the_concatenation():
    """dwg_groupnorm_apply_cs2: GroupNorm(+SiLU) of cat([x1, x2], channels) from the two tensors and their epilogue statistics,
    against the one-tensor path on the materialised concatenation (bitwise: same statistics, same arithmetic)."""
    from dwg import ops
    torch.manual_seed(11)
    for C1, C2 in ((640, 320), (1280, 640), (320, 320)):
        xin = torch.randn(2, 32, 32, 64, device=DEV).half()
        w1 = (torch.randn(C1, 3, 3, 64, device=DEV) * 0.05).half()
        w2 = (torch.randn(C2, 3, 3, 64, device=DEV) * 0.05).half()
        x1, x2 = ops.conv2d_nhwc(xin, w1, stats=True), ops.conv2d_nhwc(xin, w2, stats=True)
        g, b = torch.rand(C1 + C2, device=DEV) + 0.5, torch.randn(C1 + C2, device=DEV) * 0.1
        y = ops.group_norm_cat(x1, x2, g, b, 32, 1e-5, True)
        xc = ops.cat_channels(x1, x2)
        ref = ops.group_norm(xc, g, b, 32, 1e-5, True, colstats=xc._cs)
        assert torch.equal(y, ref)
        ref32 = torch.nn.functional.silu(torch.nn.functional.group_norm(xc.float().permute(0, 3, 1, 2), 32, g, b, 1e-5)).permute(0, 2, 3, 1)
        assert float((y.float() - ref32).abs().max()) < 2e-2
