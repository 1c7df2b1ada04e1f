"""GPU: diffusion parity AT THE BENCHMARK'S OWN SIZE AND SETTINGS: SD1.5-width UNet + ControlNet
(latent 64^2, CFG batch 2), VAE encoder forward + input gradient at 512^2, and the full guidance __call__ at
guidance_scale = 50 (configs/__init__.py:255), through dwg and through oracle/diffusion.py moved to the GPU in STRICT
fp32 (allow_tf32 = False for matmul and cuDNN).  Stated tolerances (rel-L2), measured values in brackets
(gpurun_out/r2_parity_*.txt, profiles/r2_parity.json):

    eps (each CFG row)                     <= 5e-3   [~2e-3]      SURVEY section 7 asked for <= 1e-2
    latents                                <= 5e-3   [~1e-3]
    e_c - e_u                              <= 3e-2   [~1.5e-2]
    SDS gradient at scale 50               <= 3e-2   [~1.3e-2]    = the error of (e_c - e_u): the scale multiplies signal and error alike
    dL/d(image)                            <= 3e-2   [~1.3e-2]
    the reference's own default arithmetic (TF32 cuDNN convolutions) vs strict fp32: gradient 6e-3 -- same order.

The activations are fp16 (reference option diffusion_fp16 / controlnet_fp16) with fp32 accumulation; with bf16
activations the same test measures 1.1e-2 / 8.5e-2 (eps / gradient), which is why the path is fp16.
Also asserted: two runs of the same step, and the single-stream / two-stream / prepared schedules, agree BITWISE.
"""
import pytest
import torch

from dwg.diffusion import guidance as G, weights as W

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope='module')
def sd15():
    cfg, vcfg = W.SD15, W.VAE15
    u_sd, c_sd, v_sd = W.make_unet(cfg), W.make_controlnet(cfg), W.make_vae_encoder(vcfg)
    g = torch.Generator().manual_seed(5)
    inp = {
        'img': torch.rand(1, 3, 512, 512, generator=g).to(DEV),
        'cond': (torch.rand(1, 3, 512, 512, generator=g) > 0.97).float().to(DEV),
        'emb': {'neg': torch.randn(1, 77, cfg['ctx_dim'], generator=g).to(DEV), 'text': torch.randn(1, 77, cfg['ctx_dim'], generator=g).to(DEV)},
        'noise': torch.randn(1, 4, 64, 64, generator=g).to(DEV), 'veps': torch.randn(1, 4, 64, 64, generator=g).to(DEV),
        't': torch.tensor([500], device=DEV),
    }
    gd = G.ControlNetScoreDistillation(u_sd, c_sd, v_sd, cfg, vcfg, DEV, guidance_scale=50.0)
    # ---- strict-fp32 oracle on the GPU
    from oracle import diffusion as od
    tf = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        cu = lambda sd: {k: v.to(DEV) for k, v in sd.items()}
        u, c, v = cu(u_sd), cu(c_sd), cu(v_sd)
        im = inp['img'].clone().requires_grad_(True)
        lat = od.vae_encode_latents(v, vcfg, im, inp['veps'])
        with torch.no_grad():
            ln = od.add_noise(lat.detach(), inp['noise'], inp['t'])
            ctx = torch.cat([inp['emb']['neg'], inp['emb']['text']], 0)
            x2 = torch.cat([ln] * 2, 0)
            down, mid = od.controlnet_forward(c, cfg, x2, inp['t'], ctx, inp['cond'].repeat(2, 1, 1, 1))
            eps = od.unet_forward(u, cfg, x2, inp['t'], ctx, down, mid)
            e_u, e_c = eps.chunk(2)
            grad = e_u + 50.0 * (e_c - e_u) - inp['noise']
        (lat * grad).sum().backward()
        ref = {'lat': lat.detach(), 'eps': eps, 'diff': e_c - e_u, 'grad': grad, 'gimg': im.grad.clone(), 'x2': x2, 'ctx': ctx}
        del u, c, v
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf
    torch.cuda.empty_cache()
    return gd, inp, ref


def _step(gd, inp, **kw):
    im = inp['img'].clone().requires_grad_(True)
    out = gd(im, inp['emb'], cond_inputs=inp['cond'], noise=inp['noise'], vae_eps=inp['veps'], **kw)
    out['diffusion_loss'].backward()
    torch.cuda.synchronize()
    return out, im.grad.clone()


def test_sd15_unet_controlnet_eps_vs_strict_fp32_oracle(sd15):
    gd, inp, ref = sd15
    with torch.no_grad():
        gd.timestep, gd._prepared = inp['t'], None
        eps = gd._predict(ref['x2'], ref['ctx'], inp['cond'])
    assert torch.isfinite(eps).all()
    for row in range(2):
        assert rel(eps[row], ref['eps'][row]) < 5e-3, (row, rel(eps[row], ref['eps'][row]))
    assert rel(eps[1:] - eps[:1], ref['diff']) < 3e-2, rel(eps[1:] - eps[:1], ref['diff'])


def test_sd15_vae_encode_and_input_gradient(sd15):
    gd, inp, ref = sd15
    im = inp['img'].clone().requires_grad_(True)
    lat = gd.encode_images(im, inp['veps'])
    assert rel(lat, ref['lat']) < 5e-3, rel(lat, ref['lat'])
    (lat * ref['grad']).sum().backward()
    assert rel(im.grad, ref['gimg']) < 1e-2, rel(im.grad, ref['gimg'])


def test_sd15_full_call_at_guidance_scale_50(sd15):
    gd, inp, ref = sd15
    out, gimg = _step(gd, inp, timestep=inp['t'])
    assert rel(out['latents'], ref['lat']) < 5e-3
    e_grad, e_img = rel(out['gradients'], ref['grad']), rel(gimg, ref['gimg'])
    assert e_grad < 3e-2, e_grad
    assert e_img < 3e-2, e_img
    assert float(out['diffusion_loss']) == 1.0 and out['targets'].shape == out['latents'].shape


def test_sd15_step_is_bitwise_reproducible_across_runs_and_schedules(sd15):
    gd, inp, _ = sd15
    base, gbase = _step(gd, inp, timestep=inp['t'])
    again, gagain = _step(gd, inp, timestep=inp['t'])
    assert torch.equal(again['gradients'], base['gradients']) and torch.equal(gagain, gbase)
    gd.two_streams = False
    try:
        single, gsingle = _step(gd, inp, timestep=inp['t'])
    finally:
        gd.two_streams = True
    assert torch.equal(single['gradients'], base['gradients']) and torch.equal(gsingle, gbase)
    gd.prepare(inp['emb'], inp['cond'], timestep=inp['t'])
    prep, gprep = _step(gd, inp)
    assert int(prep['timestep'][0]) == 500
    assert torch.equal(prep['gradients'], base['gradients']) and torch.equal(gprep, gbase)
