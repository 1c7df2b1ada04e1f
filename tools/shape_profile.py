"""Per-shape kernel durations of the tensor-core launches of one guidance pass (VAE fwd/bwd +
ControlNet + UNet), from CUPTI with PDL disabled (DWG_NO_PDL=1) so that a kernel's duration does
not include waiting for its predecessor.  python tools/shape_profile.py [--tiny]"""
import collections
import os
import sys

os.environ['DWG_NO_PDL'] = '1'
import torch  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'dreamwaltz-g_b200'))
from dwg import ops  # noqa: E402
from dwg.diffusion import guidance as G, weights as W  # noqa: E402


def main():
    tiny = '--tiny' in sys.argv
    dev = 'cuda:0'
    cfg, vcfg = (W.TINY, W.TINY_VAE) if tiny else (W.SD15, W.VAE15)
    g = G.ControlNetScoreDistillation(W.make_unet(cfg), W.make_controlnet(cfg), W.make_vae_encoder(vcfg), cfg, vcfg, dev, seed=1)
    gen = torch.Generator().manual_seed(7)
    emb = {'neg': torch.randn(1, 77, cfg['ctx_dim'], generator=gen).to(dev), 'text': torch.randn(1, 77, cfg['ctx_dim'], generator=gen).to(dev)}
    S = 512 if not tiny else 64
    cond = (torch.rand(1, 3, S, S, generator=gen) > 0.97).float().to(dev)

    def run():
        img = torch.rand(1, 3, S, S, device=dev, requires_grad=True)
        res = g(img, emb, cond_inputs=cond)
        res['diffusion_loss'].backward()
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    ops.PROFILE = []
    with profile(activities=[ProfilerActivity.CUDA]) as pr:
        run()
        torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    evs = [e for e in pr.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    names = ('gemm_kernel', 'fa_fwd_kernel')
    main_k = [e for e in evs if any(n in e.name for n in names)]
    assert len(main_k) == len(prof), (len(main_k), len(prof))
    # attach a splitk_finalize to the gemm that precedes it
    fin = collections.defaultdict(float)
    idx = {id(e): i for i, e in enumerate(main_k)}
    last = None
    for e in evs:
        if id(e) in idx:
            last = idx[id(e)]
        elif 'splitk_finalize' in e.name and last is not None:
            fin[last] += e.time_range.end - e.time_range.start
    agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0])
    for i, (e, (_, _, fl, kind)) in enumerate(zip(main_k, prof)):
        r = agg[kind]
        r[0] += (e.time_range.end - e.time_range.start); r[1] += fin[i]; r[2] += fl; r[3] += 1
    tot = sum(v[0] + v[1] for v in agg.values())
    print(f'{"shape":46s} {"n":>3s} {"us/call":>8s} {"fin us":>7s} {"total ms":>8s} {"%":>5s} {"TFLOP/s":>8s}')
    for kind, (us, fu, fl, n) in sorted(agg.items(), key=lambda kv: -(kv[1][0] + kv[1][1])):
        print(f'{kind:46s} {n:3d} {us / n:8.1f} {fu / n:7.1f} {(us + fu) / 1e3:8.3f} {100 * (us + fu) / tot:5.1f} {fl / ((us + fu) * 1e-6) / 1e12:8.1f}')
    print(f'total tensor-core ms {tot / 1e3:.3f}, flops {sum(v[2] for v in agg.values()) / 1e12:.3f} T, avg {sum(v[2] for v in agg.values()) / (tot * 1e-6) / 1e12:.1f} TFLOP/s')
    other = collections.defaultdict(lambda: [0.0, 0])
    for e in evs:
        if id(e) not in idx and 'splitk_finalize' not in e.name:
            other[e.name[:90]][0] += e.time_range.end - e.time_range.start
            other[e.name[:90]][1] += 1
    print('--- other kernels')
    for k, (us, n) in sorted(other.items(), key=lambda kv: -kv[1][0])[:25]:
        print(f'{k:90s} {n:4d} {us / 1e3:8.3f} ms')
    t0, t1 = evs[0].time_range.start, evs[-1].time_range.end
    print(f'span {1e-3 * (t1 - t0):.3f} ms, sum of kernels {1e-3 * sum(e.time_range.end - e.time_range.start for e in evs):.3f} ms')


if __name__ == '__main__':
    main()
