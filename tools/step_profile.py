"""Per-section and per-GEMM-shape time breakdown of one SDS step (un-graphed, CUDA events, GPU kept
busy ahead of the CPU).  python tools/step_profile.py [--tiny]"""
import collections
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'dreamwaltz-g_b200'))
import bench  # noqa: E402
from dwg import ops  # noqa: E402


def main():
    tiny = '--tiny' in sys.argv
    dev = 'cuda:0'
    sc = bench.Scene(dev, 0, tiny=tiny)
    pose, data = sc.next_view()
    pose_dev = {k: v.to(dev) for k, v in pose.items()}
    for _ in range(2):
        sc.step(pose_dev, data, sc.d_embeds, sc.d_cond)
    torch.cuda.synchronize()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    marks = []

    def mark(name):
        e = ev(); e.record(); marks.append((name, e))
    ops.PROFILE = []
    for _ in range(5):
        torch.cuda._sleep(int(4e8))
    mark('start')
    for p in sc.params:
        p.grad = None
    gs = sc.avatar.animate(pose_dev); mark('animate_fwd')
    out = sc.renderer.render(data, gs); mark('raster_fwd')
    g = sc.guidance
    img = out['image_chw'].unsqueeze(0)
    lat = g.encode_images(img); mark('vae_fwd')
    g.timestep = g.get_timestep(1)
    with torch.no_grad():
        noise = torch.randn_like(lat)
        ln = g.add_noise(lat.detach(), noise, g.timestep)
        ctx = torch.cat([sc.d_embeds['neg'], sc.d_embeds['text']], 0)
        x2 = torch.cat([ln] * 2, 0)
        cond = sc.d_cond.repeat_interleave(2, 0)
        down, mid = g.controlnet.forward(x2, g.timestep, ctx, cond); mark('controlnet')
        eps = g.unet.forward(x2, g.timestep, ctx, down, mid); mark('unet')
        e_u, e_c = eps.chunk(2)
        grad, _ = ops.sds_grad(e_u.contiguous(), e_c.contiguous(), noise, 50.0, 1.0)
    gimg, = torch.autograd.grad(lat, img, grad, retain_graph=True); mark('vae_bwd')
    img.backward(gimg); mark('raster_bwd+animate_bwd')
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    prev = marks[0][1]
    print('--- sections (ms)')
    for name, e in marks[1:]:
        print(f'{name:28s} {prev.elapsed_time(e):9.3f}')
        prev = e
    print(f"{'total':28s} {marks[0][1].elapsed_time(marks[-1][1]):9.3f}")
    agg = collections.defaultdict(lambda: [0.0, 0.0, 0])
    for a, b, f, kind in prof:
        r = agg[kind]
        r[0] += a.elapsed_time(b); r[1] += f; r[2] += 1
    print('--- tensor-core launches by shape (top 40 by time)')
    tot = sum(v[0] for v in agg.values())
    for kind, (ms, fl, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
        print(f'{kind:44s} n={n:3d} ms={ms:8.3f} ({100 * ms / tot:4.1f}%) TFLOPs={fl / (ms * 1e-3) / 1e12:7.1f}')
    print(f'total tensor-core ms {tot:.3f}, flops {sum(v[1] for v in agg.values()) / 1e12:.3f} T')
    # ---- every kernel by CUDA time (CUPTI), one more step
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as pr:
        sc.step(pose_dev, data, sc.d_embeds, sc.d_cond)
        torch.cuda.synchronize()
    print(pr.key_averages().table(sort_by='cuda_time_total', row_limit=45, max_name_column_width=70))


if __name__ == '__main__':
    main()
