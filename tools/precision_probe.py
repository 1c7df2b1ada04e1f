"""What a 16-bit pipeline can reach: the torch oracle run with bf16 / fp16 weights+activations (cuDNN/cuBLAS, fp32
accumulate) against the strict-fp32 oracle, SD1.5 sizes, guidance scale 50.  Calibrates the dwg numbers of
tools/diffusion_parity.py and decides the activation type of the tcgen05 path."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'dreamwaltz-g_b200')):
    sys.path.insert(0, p)
import torch  # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def main():
    from dwg.diffusion import weights as W
    from oracle import diffusion as od
    dev = 'cuda'
    cfg, vcfg, hw = W.SD15, W.VAE15, 512
    u_sd, c_sd, v_sd = W.make_unet(cfg), W.make_controlnet(cfg), W.make_vae_encoder(vcfg)
    g = torch.Generator().manual_seed(5)
    img = torch.rand(1, 3, hw, hw, generator=g).to(dev)
    cond = (torch.rand(1, 3, hw, hw, generator=g) > 0.97).float().to(dev)
    emb = {'neg': torch.randn(1, 77, cfg['ctx_dim'], generator=g).to(dev), 'text': torch.randn(1, 77, cfg['ctx_dim'], generator=g).to(dev)}
    noise = torch.randn(1, 4, hw // 8, hw // 8, generator=g).to(dev)
    veps = torch.randn(1, 4, hw // 8, hw // 8, generator=g).to(dev)
    t = torch.tensor([500], device=dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False

    def run(dt, grad_scale=1.0):
        cu = lambda sd: {k: v.to(dev).to(dt) for k, v in sd.items()}
        u, c, v = cu(u_sd), cu(c_sd), cu(v_sd)
        im = img.clone().requires_grad_(True)
        lat = od.vae_encode_latents(v, vcfg, im.to(dt), veps.to(dt)).float()
        with torch.no_grad():
            ln = od.add_noise(lat.detach(), noise, t)
            ctx = torch.cat([emb['neg'], emb['text']], 0)
            x2 = torch.cat([ln] * 2, 0)
            down, mid = od.controlnet_forward(c, cfg, x2.to(dt), t, ctx.to(dt), cond.repeat(2, 1, 1, 1).to(dt))
            eps = od.unet_forward(u, cfg, x2.to(dt), t, ctx.to(dt), down, mid).float()
            e_u, e_c = eps.chunk(2)
            grad = e_u + 50.0 * (e_c - e_u) - noise
        return {'lat': lat.detach(), 'eps': eps, 'diff': e_c - e_u, 'grad': grad, 'im': im, 'latg': lat}

    ref = run(torch.float32)
    (ref['latg'] * ref['grad']).sum().backward()
    out = {}
    for name, dt in (('bf16', torch.bfloat16), ('fp16', torch.float16)):
        r = run(dt)
        res = {k: rel(r[k], ref[k]) for k in ('lat', 'eps', 'diff', 'grad')}
        res['finite'] = bool(torch.isfinite(r['eps']).all())
        # image gradient for the ORACLE's latent gradient, with and without a power-of-two loss scale
        for sc in (1.0, 256.0):
            r2 = run(dt)
            (r2['latg'] * ref['grad'] * sc).sum().backward()
            res[f'gimg_scale{int(sc)}'] = rel(r2['im'].grad / sc, ref['im'].grad)
        out[name] = res
    out['gimg_abs'] = {'max': float(ref['im'].grad.abs().max()), 'mean': float(ref['im'].grad.abs().mean())}
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
