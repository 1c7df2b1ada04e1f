"""Run-to-run determinism of the diffusion blocks (same inputs, same mode, twice): bitwise comparison per component,
with optional knobs (DWG_NO_PDL=1, --ks1 forces split-K off) to localise the source of any difference."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'dreamwaltz-g_b200')):
    sys.path.insert(0, p)
import torch  # noqa: E402


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--tiny', action='store_true')
    ap.add_argument('--ks1', action='store_true')
    args = ap.parse_args()
    from dwg import ops
    from dwg._lib import lib
    from dwg.diffusion import model as M, weights as W
    dev = 'cuda'
    cfg, vcfg = (W.TINY, W.TINY_VAE) if args.tiny else (W.SD15, W.VAE15)
    hw = 128 if args.tiny else 512
    if args.ks1:
        lib().dwg_gemm_tune(0, 1)
    g = torch.Generator().manual_seed(5)
    img = torch.rand(1, 3, hw, hw, generator=g).to(dev)
    cond = (torch.rand(1, 3, hw, hw, generator=g) > 0.97).float().to(dev)
    ctx = torch.randn(2, 77, cfg['ctx_dim'], generator=g).to(dev)
    x2 = torch.randn(1, 4, hw // 8, hw // 8, generator=g).to(dev).repeat(2, 1, 1, 1)
    veps = torch.randn(1, 4, hw // 8, hw // 8, generator=g).to(dev)
    t = torch.tensor([500], device=dev)
    un, cn, vae = M.UNet(W.make_unet(cfg), cfg, dev), M.ControlNet(W.make_controlnet(cfg), cfg, dev), M.VAEEncoder(W.make_vae_encoder(vcfg), vcfg, dev)
    print('PDL', os.environ.get('DWG_NO_PDL') != '1', 'ks1', args.ks1, 'tiny', args.tiny)

    def twice(name, fn):
        a = fn(); torch.cuda.synchronize()
        b = fn(); torch.cuda.synchronize()
        a = a if isinstance(a, (list, tuple)) else [a]
        b = b if isinstance(b, (list, tuple)) else [b]
        eq = all(torch.equal(x, y) for x, y in zip(a, b))
        worst = max(rel(x, y) for x, y in zip(a, b))
        print(f'{name:34s} bitwise_equal={eq} worst_rel={worst:.3e}')
        return a

    # single layers
    xa = torch.randn(2, 64 if not args.tiny else 16, 64 if not args.tiny else 16, cfg['block_out'][0], generator=torch.Generator().manual_seed(1)).to(dev).half()
    W0 = un.W
    twice('groupnorm+silu', lambda: M.gn(W0, 'down_blocks.0.resnets.0.norm1', xa, cfg['groups'], 1e-5, True))
    twice('conv3x3', lambda: M.conv(W0, 'down_blocks.0.resnets.0.conv1', xa))
    twice('resnet', lambda: un.resnet('down_blocks.0.resnets.0', xa, None))
    un._ctx_kv = None
    twice('transformer', lambda: un.transformer('down_blocks.0.attentions.0', xa, ctx.half(), un.heads_at(0)))
    twice('vae forward', lambda: vae.forward(img, veps))
    tape = []
    lat = vae.forward(img, veps, tape)
    gl = torch.randn_like(lat)
    twice('vae backward', lambda: vae.backward(tape, gl))
    twice('controlnet', lambda: (lambda d, m: d + [m])(*cn.forward(x2, t, ctx, cond)))
    dn, md = cn.forward(x2, t, ctx, cond)
    twice('unet', lambda: un.forward(x2, t, ctx, dn, md))
    twice('unet (no residuals)', lambda: un.forward(x2, t, ctx))
    e = un.forward(x2, t, ctx, dn, md)
    print('cfg rows: |e_c - e_u| / |e| =', rel(e[1:], e[:1]))


if __name__ == '__main__':
    main()
