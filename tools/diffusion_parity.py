"""Diffusion parity at the BENCHMARK's own size and settings (SD1.5 widths, latent 64^2, CFG batch 2, VAE 512^2,
guidance scale 50): dwg (fp16 tensor-core path) against oracle/diffusion.py moved to the GPU in strict fp32
(allow_tf32 = False for matmul AND cuDNN), plus the reference's own default arithmetic (torch defaults: TF32 cuDNN
convolutions, fp32 matmuls) against the same strict-fp32 oracle as a yardstick, plus run-to-run determinism.

    gpurun -- python tools/diffusion_parity.py [--tiny] > gpurun_out/diffusion_parity.txt
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'dreamwaltz-g_b200')):
    sys.path.insert(0, p)
import torch  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--tiny', action='store_true')
    ap.add_argument('--scale', type=float, default=50.0)
    ap.add_argument('--t', type=int, default=500)
    args = ap.parse_args()
    from dwg.diffusion import guidance as G, model as M, weights as W
    from oracle import diffusion as od
    dev = 'cuda'
    cfg, vcfg = (W.TINY, W.TINY_VAE) if args.tiny else (W.SD15, W.VAE15)
    img_hw = 128 if args.tiny else 512
    lat_hw = img_hw // 8
    u_sd, c_sd, v_sd = W.make_unet(cfg), W.make_controlnet(cfg), W.make_vae_encoder(vcfg)
    cu = lambda sd: {k: v.to(dev) for k, v in sd.items()}
    u_g, c_g, v_g = cu(u_sd), cu(c_sd), cu(v_sd)
    g = torch.Generator().manual_seed(5)
    img = torch.rand(1, 3, img_hw, img_hw, generator=g).to(dev)
    cond = (torch.rand(1, 3, img_hw, img_hw, generator=g) > 0.97).float().to(dev)
    emb = {'neg': torch.randn(1, 77, cfg['ctx_dim'], generator=g).to(dev), 'text': torch.randn(1, 77, cfg['ctx_dim'], generator=g).to(dev)}
    noise = torch.randn(1, 4, lat_hw, lat_hw, generator=g).to(dev)
    veps = torch.randn(1, 4, lat_hw, lat_hw, generator=g).to(dev)
    t = torch.tensor([args.t], device=dev)
    res = {'config': 'TINY' if args.tiny else 'SD1.5 / VAE 512^2', 'guidance_scale': args.scale, 'timestep': args.t}

    def oracle_run(tf32_conv):
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = tf32_conv
        im = img.clone().requires_grad_(True)
        lat = od.vae_encode_latents(v_g, vcfg, im, veps)
        with torch.no_grad():
            ln = od.add_noise(lat.detach(), noise, t)
            ctx = torch.cat([emb['neg'], emb['text']], 0)
            x2 = torch.cat([ln] * 2, 0)
            cond2 = cond.repeat(2, 1, 1, 1)
            down, mid = od.controlnet_forward(c_g, cfg, x2, t, ctx, cond2)
            eps = od.unet_forward(u_g, cfg, x2, t, ctx, down, mid)
            e_u, e_c = eps.chunk(2)
            npred = e_u + args.scale * (e_c - e_u)
            grad = npred - noise
        (lat * grad).sum().backward()
        return {'lat': lat.detach(), 'eps': eps, 'diff': e_c - e_u, 'grad': grad, 'gimg': im.grad.clone(), 'down0': down[0], 'mid': mid,
                'x2': x2}

    ref = oracle_run(False)
    ref_tf32 = oracle_run(True)
    torch.backends.cudnn.allow_tf32 = True
    res['reference_default_tf32conv_vs_fp32'] = {k: rel(ref_tf32[k], ref[k]) for k in ('lat', 'eps', 'diff', 'grad', 'gimg')}
    res['norms'] = {'eps': float(ref['eps'].norm()), 'e_c-e_u': float(ref['diff'].norm()), 'grad': float(ref['grad'].norm()),
                    'noise': float(noise.norm()), 'lat': float(ref['lat'].norm())}

    gd = G.ControlNetScoreDistillation(u_sd, c_sd, v_sd, cfg, vcfg, dev, guidance_scale=args.scale, default_image_size=img_hw)

    def dwg_run(feed_ref_latents=False):
        im = img.clone().requires_grad_(True)
        out = gd(im, emb, cond_inputs=cond, timestep=t, noise=noise, vae_eps=veps)
        out['diffusion_loss'].backward()
        torch.cuda.synchronize()
        return out, im.grad.clone()

    out, gimg = dwg_run()
    npd = out['noise_pred']
    res['dwg_vs_fp32'] = {'lat': rel(out['latents'], ref['lat']), 'grad': rel(out['gradients'], ref['grad']), 'gimg': rel(gimg, ref['gimg'])}
    # eps of the two CFG rows on the ORACLE's noisy latents (isolates UNet+ControlNet from the VAE error)
    with torch.no_grad():
        ctx = torch.cat([emb['neg'], emb['text']], 0)
        gd.timestep = t
        gd._prepared = None
        eps = gd._predict(ref['x2'], ctx, cond)
    res['dwg_vs_fp32'].update({'eps_on_oracle_latents': rel(eps, ref['eps']), 'diff_on_oracle_latents': rel(eps[1:] - eps[:1], ref['diff']),
                               'grad_on_oracle_latents': rel(eps[:1] + args.scale * (eps[1:] - eps[:1]) - noise, ref['grad'])})
    # VAE input gradient for the ORACLE's latent gradient
    im = img.clone().requires_grad_(True)
    lat = gd.encode_images(im, veps)
    (lat * ref['grad']).sum().backward()
    res['dwg_vs_fp32']['gimg_for_oracle_grad'] = rel(im.grad, ref['gimg'])
    # determinism: same inputs, same mode, twice; and the three scheduling modes
    out2, gimg2 = dwg_run()
    res['run_to_run'] = {'grad_max_abs': float((out2['gradients'] - out['gradients']).abs().max()), 'grad_rel': rel(out2['gradients'], out['gradients']),
                         'gimg_rel': rel(gimg2, gimg), 'bitwise_equal': bool(torch.equal(out2['gradients'], out['gradients']) and torch.equal(gimg2, gimg))}
    gd.two_streams = False
    out3, gimg3 = dwg_run()
    gd.two_streams = True
    res['two_streams_vs_single'] = {'grad_rel': rel(out3['gradients'], out['gradients']), 'gimg_rel': rel(gimg3, gimg),
                                    'bitwise_equal': bool(torch.equal(out3['gradients'], out['gradients']))}
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()
